"""Batched-sweep tuning of reach_build_kernel: CTA size, resident CTAs per SM, shared sort-buffer size.
Usage: python scripts/tune_batch.py            (driver: one subprocess per configuration)
       python scripts/tune_batch.py one T B threads   (worker; knobs come from ARMOUR_TUNE_* in the environment)"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]


def worker(T, B, threads):
    import armour_b200 as ab
    if os.environ.get("ARMOUR_TUNE_LIB"): ab.LIB_PATH = os.path.join(ab.PKG_DIR, os.environ["ARMOUR_TUNE_LIB"])
    from problems import make_problem
    n_obs = 20
    pb = ab.Planner(T=T, max_obstacles=n_obs, device=0, batch=B, threads_per_cta=threads)
    bp = [make_problem(5000 + i, n_obs) for i in range(B)]
    pb.upload_problems(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]),
                       np.concatenate([q[4] for q in bp]), n_obs)
    pb.build_resident()
    ms = []
    for _ in range(3):
        pb.build_resident()
        ms.append(pb.last_build_ms()[0])
    print("RESULT %.3f ms per batch, %.1f builds/s" % (np.mean(ms), B * 1e3 / np.mean(ms)), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        return worker(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    for B in (64,):
        for threads, minb, scap, tcap in ((128, 4, 1024, 256), (128, 4, 1408, 300), (128, 4, 768, 128), (128, 3, 2048, 512), (128, 6, 512, 256), (256, 2, 2048, 512), (256, 2, 1024, 256)):
            env = dict(os.environ, ARMOUR_TUNE_MINB=str(minb), ARMOUR_TUNE_SCAP=str(scap), ARMOUR_TUNE_TCAP=str(tcap))
            out = subprocess.run([sys.executable, __file__, "one", "128", str(B), str(threads)], env=env, capture_output=True, text=True)
            res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
            print("B=%d threads=%d minb=%d scap=%d tcap=%d:" % (B, threads, minb, scap, tcap), res[0] if res else "FAILED " + out.stderr[-200:], flush=True)


if __name__ == "__main__":
    main()
