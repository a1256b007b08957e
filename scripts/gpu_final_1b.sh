#!/bin/bash
# part b: full capture of the sweep-shaped reach kernel (16 problems x 128 intervals) and the per-phase cycle table
mkdir -p gpurun_out
python scripts/tune_sweep.py one 16 10 > gpurun_out/f1_sweep_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:reach_build -s 1 -c 1 -o gpurun_out/r2_final_sweep python scripts/tune_sweep.py one 16 10 > gpurun_out/f1_ncu3.log 2>&1
du -sh gpurun_out
