"""How often could a cached sort order be reused between consecutive time intervals of one problem?
Measurement build (make variant VARIANT=keyhash EXTRA=-DARMOUR_KEYHASH): every sorting operation records a hash of its result's
key list and its candidate count N per (interval, operation number).  An operation of interval s can reuse the order of interval
s-1 only if BOTH operand key lists are unchanged; operands are results of earlier operations (or of stage A), so "result key list
unchanged" per operation bounds the reuse rate from above.  Reported: per-operation equality rate between consecutive intervals,
unweighted and weighted by candidate count, and the rate for 'all results so far unchanged' (what a chain of dependent operations needs)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
ab.LIB_PATH = os.path.join(ab.PKG_DIR, "libarmour_b200_keyhash.so")
from problems import make_problem
T, B, n_obs, OPS = 128, 2, 4, 512
for seed in (5000, 5003):
    pb = ab.Planner(T=T, max_obstacles=n_obs, device=0, batch=B)
    bp = [make_problem(seed + i, n_obs) for i in range(B)]
    pb.build_batch(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]), np.concatenate([q[4] for q in bp]), n_obs)
    out = np.zeros((B * T, OPS, 2), dtype=np.uint64)
    rc = pb.L.armour_debug_keyhash(out.ctypes.data_as(C.POINTER(C.c_ulonglong)), C.c_int(B * T))
    assert rc == 0
    for p in range(B):
        h = out[p * T:(p + 1) * T, :, 0]
        n = out[p * T:(p + 1) * T, :, 1].astype(np.float64)
        nops = int((n.sum(axis=0) > 0).nonzero()[0].max()) + 1
        h, n = h[:, :nops], n[:, :nops]
        same = h[1:] == h[:-1]                                   # [T-1, nops]
        prefix_same = np.logical_and.accumulate(same, axis=1)    # every result up to this operation unchanged
        w = n[1:]
        print("seed %d: %d sorting operations per interval, %.0f candidates per interval" % (seed + p, nops, n.sum() / T))
        print("   result key list identical to the previous interval's: %.1f %% of operations, %.1f %% of candidates" % (100 * same.mean(), 100 * (same * w).sum() / w.sum()))
        print("   ... and every earlier result of the interval identical too: %.1f %% of operations, %.1f %% of candidates" % (100 * prefix_same.mean(), 100 * (prefix_same * w).sum() / w.sum()))
        first = np.where(same.all(axis=1), nops, np.argmin(same, axis=1))
        print("   first differing operation per interval pair: median %d of %d (quartiles %d / %d)" % (np.median(first), nops, np.percentile(first, 25), np.percentile(first, 75)))
    pb.close()
