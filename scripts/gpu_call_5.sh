#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
tail -n 25 gpurun_out/c5_pytest.log
timeout 300 python scripts/eval_latency.py 20 2>&1 | tee gpurun_out/c5_eval_latency_20.log | tail -n 3
python scripts/profile_step.py > gpurun_out/c5_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5_launches.csv python scripts/profile_step.py > gpurun_out/c5_ncu.log 2>&1
grep -v "^==" gpurun_out/c5_launches.csv | cut -d, -f5,12- | tail -n 12
