#!/bin/bash
mkdir -p gpurun_out
for lib in libarmour_b200.so libarmour_b200_nosharedfk.so libarmour_b200.so libarmour_b200_nosharedfk.so; do
  echo "single $lib: $(ARMOUR_TUNE_LIB=$lib timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/c20_sharedfk.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -n 5
