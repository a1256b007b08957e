"""Load imbalance of the per-key segment walk (measurement build: make variant VARIANT=segstat EXTRA=-DARMOUR_SEGSTAT).
Today a warp iteration of the walk lasts as long as its longest segment (the head lane computes every term of its segment one
after the other, the other lanes of the segment idle).  If every lane computed its own term and the heads only added them up in
order, an iteration would cost one term plus a short chain of additions.  Reported: mean candidates, heads and longest segment
per warp iteration — the ratio 'longest segment : 1' bounds what balancing the walk could save."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
ab.LIB_PATH = os.path.join(ab.PKG_DIR, "libarmour_b200_segstat.so")
from problems import make_problem
out = (C.c_ulonglong * 4)()
def report(label, L):
    L.armour_debug_segstat(out, 1)
    it, cand, heads, mx = [float(v) for v in out]
    print("%-34s warp iterations %9.0f: candidates %.1f, heads %.1f, longest segment %.2f  (mean segment %.2f)" % (label, it, cand / it, heads / it, mx / it, cand / max(heads, 1)))
p = ab.Planner(T=128, max_obstacles=20)
q0, qd0, qdd0, _, obs = make_problem(100000, 20)
p.build(q0, qd0, qdd0, obs); p.L.armour_debug_segstat(out, 1)
p.build(q0, qd0, qdd0, obs); report("one plan (2 x 256 threads)", p.L)
B = 16
pb = ab.Planner(T=128, max_obstacles=10, batch=B)
bp = [make_problem(5000 + i, 10) for i in range(B)]
args = [np.concatenate([q[k] for q in bp]) for k in (0, 1, 2, 4)]
pb.build_batch(*args, 10); pb.L.armour_debug_segstat(out, 1)
pb.build_batch(*args, 10); report("sweep shape (128 threads), 16 problems", pb.L)
