"""Occupancy sweep of controller_update_kernel (tuning knobs ARMOUR_TUNE_CTRL_THREADS / ARMOUR_TUNE_CTRL_SMEM)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for threads, smem in ((128, 0), (128, 57000), (128, 75000), (128, 114000), (128, 200000), (64, 0), (64, 28000), (64, 37000), (64, 57000), (64, 114000), (32, 14000), (32, 28000), (256, 0), (256, 114000)):
    env = dict(os.environ, ARMOUR_TUNE_CTRL_THREADS=str(threads), ARMOUR_TUNE_CTRL_SMEM=str(smem))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_controller.py"), "262144"], env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(threads, smem, "kernel_ms %.3f" % j["kernel_ms"], flush=True)
    except Exception as e:
        print(threads, smem, "failed", out.stderr[-300:], flush=True)
