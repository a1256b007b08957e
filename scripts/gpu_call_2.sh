#!/bin/bash
# arena slot stride (max_monomials) vs L1 set conflicts: power-of-two strides against skewed ones
mkdir -p gpurun_out
for m in 1024 1032 1040 1048 1096 528 536 776; do
  echo "MCAP=$m single: $(ARMOUR_TUNE_MCAP=$m timeout 120 python scripts/quick_reach_ms.py 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/c2_mcap_single.log
python - <<'PY' 2>&1 | tee gpurun_out/c2_mcap_sweep.log
import os, subprocess, sys, json
for m in (1024, 1032, 1040, 1048, 1096, 528, 536, 776):
    env = dict(os.environ, ARMOUR_TUNE_NT="128", ARMOUR_TUNE_MINB="4", ARMOUR_TUNE_SCAP="1408", ARMOUR_TUNE_TCAP="300", ARMOUR_TUNE_MCAP=str(m))
    out = subprocess.run([sys.executable, "scripts/tune_sweep.py", "one", "128", "10"], env=env, capture_output=True, text=True, timeout=600)
    res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    print("MCAP=%d sweep:" % m, res[0] if res else "FAILED " + out.stderr[-300:], flush=True)
PY
