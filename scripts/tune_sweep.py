"""Sweep-shape tuning of reach_build_kernel (round 2): threads per interval, resident CTAs per SM, shared sort-buffer sizes,
monomial capacities.  Every configuration is checked against the first one: monomial keys and coefficients of sampled
torque / link tables must be bit-identical (digest), then timed (reach kernel only, CUDA events, 3 launches after a warm-up).

Usage: python scripts/tune_sweep.py [B] [n_obs]          driver: one subprocess per configuration, one line per result
       python scripts/tune_sweep.py one B n_obs           worker; knobs come from ARMOUR_TUNE_* in the environment"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]

CONFIGS = [
    # (threads per interval, CTAs/SM bound, shared sort entries, shared staging entries, monomial capacity)
    (128, 4, 1408, 300, 1024),   # round-1 sweep shape (reference digest)
    (128, 3, 1408, 300, 1024),   # round-2 default
    (128, 5, 1408, 300, 1024),
    (128, 2, 1408, 300, 1024),
    (128, 3, 1408, 640, 1024),
    (64, 8, 768, 192, 512),
    (64, 12, 512, 96, 512),
    (32, 12, 512, 128, 512),
    (32, 16, 384, 64, 512),
    (32, 24, 256, 32, 512),
]


def worker(B, n_obs):
    import armour_b200 as ab
    if os.environ.get("ARMOUR_TUNE_LIB"):
        ab.LIB_PATH = os.path.join(ab.PKG_DIR, os.environ["ARMOUR_TUNE_LIB"])
    from problems import make_problem
    T = 128
    pb = ab.Planner(T=T, max_obstacles=n_obs, device=0, batch=B)
    bp = [make_problem(5000 + i, n_obs) for i in range(B)]
    pb.upload_problems(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]),
                       np.concatenate([q[4] for q in bp]), n_obs)
    pb.build_resident()
    ms = []
    for _ in range(3):
        pb.build_resident()
        ms.append(pb.last_build_ms()[1])
    h = hashlib.sha1()
    for prob in sorted({0, B // 2, B - 1}):
        pb.select_problem(prob)
        for s in (0, 37, 64, 100, 127):
            for j in range(7):
                for name in ("u_nom", "links"):
                    z = pb.get_pz(name, j, s)
                    h.update(np.ascontiguousarray(z["keys"]).tobytes())
                    h.update(np.ascontiguousarray(z["coeffs"]).tobytes())
                    h.update(np.ascontiguousarray(z["center"]).tobytes())
    tr = pb.torque_radius()
    print("RESULT " + json.dumps({"ms": float(np.mean(ms)), "ms_min": float(np.min(ms)), "builds_per_s": B * 1e3 / float(np.mean(ms)),
                                  "digest": h.hexdigest()[:16], "torque_radius_sum": float(tr.sum())}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        return worker(int(sys.argv[2]), int(sys.argv[3]))
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_obs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    ref = None
    for nt, minb, scap, tcap, mcap in CONFIGS:
        env = dict(os.environ, ARMOUR_TUNE_NT=str(nt), ARMOUR_TUNE_MINB=str(minb), ARMOUR_TUNE_SCAP=str(scap), ARMOUR_TUNE_TCAP=str(tcap), ARMOUR_TUNE_MCAP=str(mcap))
        out = subprocess.run([sys.executable, __file__, "one", str(B), str(n_obs)], env=env, capture_output=True, text=True, timeout=600)
        res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        tag = "B=%d nt=%d minb=%d scap=%d tcap=%d mcap=%d" % (B, nt, minb, scap, tcap, mcap)
        if not res:
            print(tag, "FAILED", out.stderr[-400:].replace("\n", " | "), flush=True)
            continue
        r = json.loads(res[0][7:])
        if ref is None:
            ref = r
        ok = r["digest"] == ref["digest"] and abs(r["torque_radius_sum"] - ref["torque_radius_sum"]) <= 1e-9 * abs(ref["torque_radius_sum"])
        print(tag, "%.2f ms (min %.2f)  %.0f builds/s  %s" % (r["ms"], r["ms_min"], r["builds_per_s"], "parity-ok" if ok else "PARITY-MISMATCH " + r["digest"]), flush=True)


if __name__ == "__main__":
    main()
