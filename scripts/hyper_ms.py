"""hyperplane_kernel time (CUDA events) for one plan (T = 128, 20 obstacles) and for a batch of 256 problems with 10 obstacles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
if os.environ.get("ARMOUR_TUNE_LIB"): ab.LIB_PATH = os.path.join(ab.PKG_DIR, os.environ["ARMOUR_TUNE_LIB"])
from problems import make_problem
p = ab.Planner(T=128, max_obstacles=40)
for n_obs in (20, 10, 40):
    ms = []
    for s in range(6):
        q0, qd0, qdd0, _, obs = make_problem(100000 + s, n_obs)
        p.build(q0, qd0, qdd0, obs); ms.append(p.last_build_ms()[2] * 1e3)
    print("one plan, %d obstacles: hyperplane_kernel us" % n_obs, " ".join("%.1f" % m for m in ms[1:]))
B = 256
pb = ab.Planner(T=128, max_obstacles=10, batch=B)
bp = [make_problem(5000 + i, 10) for i in range(B)]
args = [np.concatenate([q[k] for q in bp]) for k in (0, 1, 2, 4)]
for _ in range(3):
    pb.build_batch(*args, 10)
print("batch of 256, 10 obstacles: hyperplane_kernel ms %.3f (reach %.2f)" % (pb.last_build_ms()[2], pb.last_build_ms()[1]))
