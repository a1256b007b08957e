"""Reach-kernel time of five single-plan builds (T = 128, 20 obstacles); ARMOUR_TUNE_LIB selects a tuning build of the library."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import armour_b200 as ab
if os.environ.get("ARMOUR_TUNE_LIB"): ab.LIB_PATH = os.path.join(ab.PKG_DIR, os.environ["ARMOUR_TUNE_LIB"])
from problems import make_problem
p = ab.Planner(T=128, max_obstacles=20, device=0)
ms=[]
for s in range(6):
    q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
    p.build(q0, qd0, qdd0, obs); ms.append(p.last_build_ms()[1])
print("reach_ms", " ".join("%.3f"%m for m in ms[1:]))
