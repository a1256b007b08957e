"""Per-phase cycle accounting of reach_build_kernel (profiling build libarmour_b200_phase.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
ab.LIB_PATH = os.path.join(ab.PKG_DIR, "libarmour_b200_phase.so")
from problems import make_problem
p = ab.Planner(T=128, max_obstacles=20)
names = ["fill", "sort level", "segment walk", "scan+compact", "element-wise", "stage A", "export", "other"]
for s in range(3):
    q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
    cyc = np.zeros(8, dtype=np.uint64); calls = np.zeros(8, dtype=np.uint64)
    p.L.armour_debug_phase_cycles(cyc.ctypes.data_as(C.POINTER(C.c_uint64)), calls.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
    p.build(q0, qd0, qdd0, obs)
    p.L.armour_debug_phase_cycles(cyc.ctypes.data_as(C.POINTER(C.c_uint64)), calls.ctypes.data_as(C.POINTER(C.c_uint64)), 0)
print("reach kernel ms", p.last_build_ms()[1], "(instrumented)")
tot = cyc.sum()
for n, c, k in zip(names, cyc, calls):
    print("%-14s %5.1f%%  calls/CTA %7.1f  cycles/call %8.1f" % (n, 100.0 * c / tot, k / 128.0, c / max(k, 1)))
print("cycles per CTA", tot / 128.0)
