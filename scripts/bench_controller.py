"""Robust-controller row: device ticks/s (resident and through the C ABI with host buffers) against the
CPU oracle on the box's host cores.  Usage: python scripts/bench_controller.py [count]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import armour_b200 as ab
if os.environ.get("ARMOUR_TUNE_LIB"):   # tuning builds
    ab.LIB_PATH = os.path.join(ab.PKG_DIR, os.environ["ARMOUR_TUNE_LIB"])
from armour_b200.controller import RobustController
import _oracle
from test_controller import states, MODEL, KR, ALPHA, V_MAX, R_THR


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    st = states(7, count)
    c = RobustController(MODEL, 0.03, device=0)
    c.upload(*st)
    for _ in range(3):
        c.update_resident(KR, ALPHA, V_MAX, R_THR)
    ms = []
    for _ in range(10):
        c.update_resident(KR, ALPHA, V_MAX, R_THR)
        ms.append(c.last_ms())
    kernel_ms = float(np.median(ms))
    out = tuple(np.zeros((count, 7)) for _ in range(3))   # reused every call, like a closed-loop sweep would
    c.update(KR, ALPHA, V_MAX, R_THR, *st, out=out)
    t0 = time.perf_counter()
    for _ in range(3):
        c.update(KR, ALPHA, V_MAX, R_THR, *st, out=out)
    e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    c.release_host_buffers()
    lat = []
    one = [a[0] for a in st]
    for _ in range(200):
        t0 = time.perf_counter()
        c.update(KR, ALPHA, V_MAX, R_THR, *one)
        lat.append((time.perf_counter() - t0) * 1e6)
    cores = len(os.sched_getaffinity(0))
    o = _oracle.OracleController(MODEL, num_threads=cores)
    sample = min(count, 20000 * cores)
    sub = [a[:sample] for a in st]
    o.update(KR, ALPHA, V_MAX, R_THR, *[a[:1000] for a in st])
    t0 = time.perf_counter()
    o.update(KR, ALPHA, V_MAX, R_THR, *sub)
    cpu_s = time.perf_counter() - t0
    peak = ab.measure_fp64_peak(0)
    # directed-rounding operation count per tick: nominal + interval + M r passes
    flop_per_tick = 5.6e4
    out = {"metric": "controller_ticks_per_s", "count": count, "kernel_ms": kernel_ms, "value": count / kernel_ms * 1e3,
           "e2e_ms": e2e_ms, "e2e": count / e2e_ms * 1e3, "single_tick_latency_us_p50": float(np.median(lat)),
           "cpu_baseline": {"value": sample / cpu_s, "cores": cores, "kind": "port", "sample": "%d ticks" % sample},
           "roofline": {"bound": "fp64", "achieved": flop_per_tick * count / kernel_ms * 1e3 / 1e12, "peak": peak, "unit": "TFLOP/s"}}
    out["roofline"]["frac"] = out["roofline"]["achieved"] / peak
    print(json.dumps(out))


if __name__ == "__main__":
    main()
