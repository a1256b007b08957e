"""Tuning sweep: single-plan latency and batched throughput for kernel variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
from problems import make_problem
B = int(os.environ.get("B", "64"))
for nt, minb, ncap in ((256, 1, 4096), (256, 2, 4096), (128, 2, 4096), (128, 4, 4096), (512, 1, 4096), (128, 4, 3072), (256, 2, 3072)):
    os.environ["ARMOUR_TUNE_MINB"] = str(minb)
    try:
        p = ab.Planner(T=128, max_obstacles=20, threads_per_cta=nt, max_entries=ncap)
        ts = []
        for s in range(4):
            q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
            p.build(q0, qd0, qdd0, obs); ts.append(p.last_build_ms()[1])
        p.close()
        pb = ab.Planner(T=128, max_obstacles=20, threads_per_cta=nt, batch=B, max_entries=ncap)
        bp = [make_problem(105000 + i, 20) for i in range(B)]
        pb.upload_problems(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]), np.concatenate([q[4] for q in bp]), 20)
        pb.build_resident(); pb.build_resident()
        tb = pb.last_build_ms()[1]
        pb.close()
        print("nt %d minb %d ncap %d: single %.3f ms | batch %d: %.2f ms -> %.0f builds/s" % (nt, minb, ncap, min(ts[1:]), B, tb, B * 1e3 / tb), flush=True)
    except Exception as e:
        print("nt %d minb %d ncap %d: %s" % (nt, minb, ncap, e), flush=True)
