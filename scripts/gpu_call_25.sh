#!/bin/bash
export ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=3 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300
for lib in libarmour_b200.so libarmour_b200_dup.so libarmour_b200.so libarmour_b200_dup.so; do
  echo "sweep $lib: $(ARMOUR_TUNE_LIB=$lib timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1 | cut -c1-100)"
done
unset ARMOUR_TUNE_NT ARMOUR_TUNE_MINB ARMOUR_TUNE_SCAP ARMOUR_TUNE_TCAP
for lib in libarmour_b200.so libarmour_b200_dup.so; do
  echo "single $lib: $(ARMOUR_TUNE_LIB=$lib timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
done
