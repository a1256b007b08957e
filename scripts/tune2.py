import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
from problems import make_problem
B = 64
bp = [make_problem(105000 + i, 20) for i in range(B)]
for groups, minb in ((2, 1), (1, 2), (1, 1)):
    os.environ["ARMOUR_TUNE_GROUPS"] = str(groups); os.environ["ARMOUR_TUNE_MINB"] = str(minb)
    p = ab.Planner(T=128, max_obstacles=20)
    ts = []
    for s in range(5):
        q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
        p.build(q0, qd0, qdd0, obs); ts.append(p.last_build_ms()[1])
    p.close()
    pb = ab.Planner(T=128, max_obstacles=20, batch=B)
    pb.upload_problems(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]), np.concatenate([q[4] for q in bp]), 20)
    pb.build_resident(); pb.build_resident()
    tb = pb.last_build_ms()[1]
    pb.close()
    print("groups %d minb %d: single %.3f ms (min of 4) | batch %d: %.2f ms -> %.0f builds/s" % (groups, minb, min(ts[1:]), B, tb, B * 1e3 / tb), flush=True)
