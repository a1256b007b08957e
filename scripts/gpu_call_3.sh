#!/bin/bash
# ncu --set full of the sweep-shaped reach kernel: 128 threads x 4 CTAs/SM against 32 threads x 16 CTAs/SM (16 problems)
mkdir -p gpurun_out
export ARMOUR_TUNE_MCAP=512
ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=4 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300 python scripts/tune_sweep.py one 16 10 > gpurun_out/c3_plain128.log 2>&1 &&
ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=4 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300 ncu --set full --clock-control none --import-source on -k regex:reach_build -s 1 -c 1 -o gpurun_out/r2_nt128 python scripts/tune_sweep.py one 16 10 > gpurun_out/c3_ncu128.log 2>&1
ARMOUR_TUNE_NT=32 ARMOUR_TUNE_MINB=16 ARMOUR_TUNE_SCAP=384 ARMOUR_TUNE_TCAP=64 python scripts/tune_sweep.py one 16 10 > gpurun_out/c3_plain32.log 2>&1 &&
ARMOUR_TUNE_NT=32 ARMOUR_TUNE_MINB=16 ARMOUR_TUNE_SCAP=384 ARMOUR_TUNE_TCAP=64 ncu --set full --clock-control none --import-source on -k regex:reach_build -s 1 -c 1 -o gpurun_out/r2_nt32 python scripts/tune_sweep.py one 16 10 > gpurun_out/c3_ncu32.log 2>&1
cat gpurun_out/c3_plain128.log gpurun_out/c3_plain32.log; tail -3 gpurun_out/c3_ncu128.log gpurun_out/c3_ncu32.log; ls -la gpurun_out/*.ncu-rep
