#!/bin/bash
mkdir -p gpurun_out
for b in 0 1 2 3 4 5 7; do echo "BPS=$b $(ARMOUR_TUNE_EVAL_BPS=$b timeout 300 python scripts/eval_latency.py 20 2>&1 | tail -n 1 | cut -c1-420)"; done 2>&1 | tee gpurun_out/c6_bps.log
