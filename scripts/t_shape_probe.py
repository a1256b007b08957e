import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
from problems import make_problem
T = int(sys.argv[1])
p = ab.Planner(T=T, max_obstacles=20, mass_uncertainty=0.05, inertia_uncertainty=0.05)
ms = []
for s in range(7):
    q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
    p.build(q0, qd0, qdd0, obs); ms.append(p.last_build_ms()[1])
print("T=%d NT=%s MINB=%s GROUPS=%s reach_ms %s" % (T, os.environ.get("ARMOUR_TUNE_NT"), os.environ.get("ARMOUR_TUNE_MINB"), os.environ.get("ARMOUR_TUNE_GROUPS"), " ".join("%.3f" % m for m in ms[2:])))
