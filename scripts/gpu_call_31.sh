#!/bin/bash
mkdir -p gpurun_out
export ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=4 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300
python scripts/tune_sweep.py one 16 10 > gpurun_out/c31_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:reach_build -s 1 -c 1 -o gpurun_out/r2_op3 python scripts/tune_sweep.py one 16 10 > gpurun_out/c31_ncu.log 2>&1
cat gpurun_out/c31_plain.log
