#!/bin/bash
# A/B of a tuning build (libarmour_b200_$1.so) against the product library: single plan, sweep, result digest
for lib in "" "libarmour_b200_$1.so"; do
  echo "== ${lib:-product}"
  echo "single: $(ARMOUR_TUNE_LIB=$lib timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
  echo "sweep: $(ARMOUR_TUNE_LIB=$lib timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1 | cut -c1-135)"
done
