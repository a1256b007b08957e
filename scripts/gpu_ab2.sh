#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -n 3
echo "single: $(timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
for cfg in "4 1408 300" "4 1536 300" "4 1664 256" "3 1408 300"; do set -- $cfg
  echo "sweep minb=$1 scap=$2 tcap=$3: $(ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=$1 ARMOUR_TUNE_SCAP=$2 ARMOUR_TUNE_TCAP=$3 timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1 | cut -c1-135)"
done
