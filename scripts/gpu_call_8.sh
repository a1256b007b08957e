#!/bin/bash
mkdir -p gpurun_out
for w in 0 1 2; do echo "HOST_WRITE=$w $(ARMOUR_TUNE_HOST_WRITE=$w timeout 300 python scripts/eval_latency.py 20 2>&1 | tail -n 1 | cut -c1-520)"; done 2>&1 | tee gpurun_out/c8_hostwrite.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -n 3
