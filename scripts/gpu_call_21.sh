#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_eval.py > gpurun_out/c21_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:constraint_eval -s 2 -c 1 -o gpurun_out/r2_eval python scripts/profile_eval.py > gpurun_out/c21_ncu.log 2>&1
cat gpurun_out/c21_plain.log
