"""Short program for ncu: one build (T = 128, 20 obstacles) and a few device-resident constraint evaluations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
from problems import make_problem
p = ab.Planner(T=128, max_obstacles=20)
q0, qd0, qdd0, _, obs = make_problem(100000, 20)
p.build(q0, qd0, qdd0, obs)
for s in range(4):
    p.upload_x(np.random.default_rng(s).uniform(-1, 1, 7))
    p.eval_resident(None)
print("ok", p.last_eval_ms())
