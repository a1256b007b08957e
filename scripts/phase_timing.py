"""Per-phase cycle accounting of reach_build_kernel (profiling build libarmour_b200_phase.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
ab.LIB_PATH = os.path.join(ab.PKG_DIR, "libarmour_b200_phase.so")
from problems import make_problem
names = ["fill", "sort level", "segment walk", "scan+compact", "element-wise", "stage A", "export", "other"]
u64p = C.POINTER(C.c_uint64)
cyc = np.zeros(8, dtype=np.uint64); calls = np.zeros(8, dtype=np.uint64)
if len(sys.argv) > 1 and sys.argv[1] == "sweep":      # the sweep shape: 64 problems x 128 intervals, 128-thread CTAs
    B, n_obs = 64, 10
    p = ab.Planner(T=128, max_obstacles=n_obs, batch=B)
    bp = [make_problem(5000 + i, n_obs) for i in range(B)]
    args = [np.concatenate([q[k] for q in bp]) for k in (0, 1, 2, 4)]
    p.build_batch(*args, n_obs)
    p.L.armour_debug_phase_cycles(cyc.ctypes.data_as(u64p), calls.ctypes.data_as(u64p), 1)
    p.build_batch(*args, n_obs)
    p.L.armour_debug_phase_cycles(cyc.ctypes.data_as(u64p), calls.ctypes.data_as(u64p), 0)
    n_cta = B * 128.0
else:
    p = ab.Planner(T=128, max_obstacles=20)
    for s in range(3):
        q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
        p.L.armour_debug_phase_cycles(cyc.ctypes.data_as(u64p), calls.ctypes.data_as(u64p), 1)
        p.build(q0, qd0, qdd0, obs)
        p.L.armour_debug_phase_cycles(cyc.ctypes.data_as(u64p), calls.ctypes.data_as(u64p), 0)
    n_cta = 128.0
print("reach kernel ms", p.last_build_ms()[1], "(instrumented)")
tot = cyc.sum()
for n, c, k in zip(names, cyc, calls):
    print("%-14s %5.1f%%  calls/interval %7.1f  cycles/call %8.1f" % (n, 100.0 * c / tot, k / n_cta, c / max(k, 1)))
print("cycles per interval", tot / n_cta)
