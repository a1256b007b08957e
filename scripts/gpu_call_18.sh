#!/bin/bash
mkdir -p gpurun_out
export ARMOUR_TUNE_NO_STRUCTURED=1 ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=4
for cfg in "1408 300" "1408 640" "1408 900" "1024 768" "1152 1024" "2048 300"; do set -- $cfg
  echo "sweep scap=$1 tcap=$2: $(ARMOUR_TUNE_SCAP=$1 ARMOUR_TUNE_TCAP=$2 timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1 | cut -c1-90)"
done 2>&1 | tee gpurun_out/c18_tcap.log
unset ARMOUR_TUNE_NT ARMOUR_TUNE_MINB ARMOUR_TUNE_NO_STRUCTURED
for cfg in "2048 512" "2048 1024" "2048 1536" "1536 1024" "3072 512"; do set -- $cfg
  echo "single scap=$1 tcap=$2: $(ARMOUR_TUNE_SCAP=$1 ARMOUR_TUNE_TCAP=$2 timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
done 2>&1 | tee -a gpurun_out/c18_tcap.log
