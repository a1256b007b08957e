#!/bin/bash
# round-2 GPU call 1: sanity tests, sweep-shape tuning, sanitizer attempt
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 120 python scripts/quick_reach_ms.py > gpurun_out/c1_single.log 2>&1
timeout 1200 python scripts/tune_sweep.py 256 10 > gpurun_out/c1_tune_sweep.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_target.py > gpurun_out/c1_sanitizer_$tool.txt 2>&1; echo "rc=$?" >> gpurun_out/c1_sanitizer_$tool.txt
done
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_single.log; cat gpurun_out/c1_tune_sweep.log; tail -5 gpurun_out/c1_sanitizer_*.txt
