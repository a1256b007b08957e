#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 15
for ns in 1 0; do
  echo "NO_STRUCTURED=$ns single: $(ARMOUR_TUNE_NO_STRUCTURED=$ns timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
  ARMOUR_TUNE_NO_STRUCTURED=$ns ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=4 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300 timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1
  ARMOUR_TUNE_NO_STRUCTURED=$ns ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=3 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300 timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1
done 2>&1 | tee gpurun_out/c15_structured.log
