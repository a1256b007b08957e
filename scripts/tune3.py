import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
from problems import make_problem
for groups, scap, tcap in ((2, 2048, 1024), (2, 2048, 512), (2, 1024, 512), (2, 1024, 256), (2, 3072, 1024), (1, 2048, 1024), (1, 4096, 1024), (1, 1024, 512)):
    os.environ["ARMOUR_TUNE_GROUPS"] = str(groups); os.environ["ARMOUR_TUNE_SCAP"] = str(scap); os.environ["ARMOUR_TUNE_TCAP"] = str(tcap)
    p = ab.Planner(T=128, max_obstacles=20)
    ts = []
    for s in range(6):
        q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
        p.build(q0, qd0, qdd0, obs); ts.append(p.last_build_ms()[1])
    p.close()
    print("groups %d scap %d tcap %d: single min %.3f median %.3f ms" % (groups, scap, tcap, min(ts[1:]), np.median(ts[1:])), flush=True)
