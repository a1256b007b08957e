#!/bin/bash
mkdir -p gpurun_out
for b in 0 3; do for f in 0 1; do echo "BPS=$b FLAG_IN_KERNEL=$f $(ARMOUR_TUNE_FLAG_IN_KERNEL=$f ARMOUR_TUNE_EVAL_BPS=$b timeout 300 python scripts/eval_latency.py 20 2>&1 | tail -n 1 | cut -c1-330)"; done; done 2>&1 | tee gpurun_out/c7_flag.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -n 3
