"""Short program for ncu: one warm-up step and two measured steps of bench.py's workload (T=128, 20 obstacles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
from problems import make_problem
p = ab.Planner(T=128, max_obstacles=20, threads_per_cta=int(os.environ.get("NT", "0")))
for s in range(3):
    q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
    p.build(q0, qd0, qdd0, obs)
    g, J = p.eval_g_jac(np.random.default_rng(s).uniform(-1, 1, 7))
print("ok", p.last_build_ms(), p.last_eval_ms(), p.kernel_launches())
