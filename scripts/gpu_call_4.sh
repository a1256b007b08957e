#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -n 25 gpurun_out/c4_pytest.log
timeout 300 python scripts/eval_latency.py 20 2>&1 | tee gpurun_out/c4_eval_latency_20.log | tail -n 3
timeout 300 python scripts/eval_latency.py 10 2>&1 | tee gpurun_out/c4_eval_latency_10.log | tail -n 3
timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -n 2
