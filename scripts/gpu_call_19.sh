#!/bin/bash
mkdir -p gpurun_out
export ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=4 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300
for lib in libarmour_b200.so libarmour_b200_nostruct.so libarmour_b200_nostruct_nodeg.so; do
  echo "sweep $lib: $(ARMOUR_TUNE_LIB=$lib ARMOUR_TUNE_NO_STRUCTURED=1 timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1 | cut -c1-90)"
done 2>&1 | tee gpurun_out/c19_variants.log
unset ARMOUR_TUNE_NT ARMOUR_TUNE_MINB ARMOUR_TUNE_SCAP ARMOUR_TUNE_TCAP
for lib in libarmour_b200.so libarmour_b200_nostruct.so libarmour_b200_nostruct_nodeg.so; do
  echo "single $lib: $(ARMOUR_TUNE_LIB=$lib timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
done 2>&1 | tee -a gpurun_out/c19_variants.log
