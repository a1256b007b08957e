#!/bin/bash
# A/B of the library under test: parity (primitives + pipeline), single plan, sweep (128 x 3) with digest
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -n 4
echo "single: $(timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=3 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300 timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1
ARMOUR_TUNE_NT=128 ARMOUR_TUNE_MINB=4 ARMOUR_TUNE_SCAP=1408 ARMOUR_TUNE_TCAP=300 timeout 300 python scripts/tune_sweep.py one 256 10 2>&1 | tail -n 1
