"""Per-iteration constraint latency (BASELINE.json configs[1]: T = 128, 20 obstacles): fused eval_g + eval_jac_g through the
C ABI with host buffers (p50 / p99 wall-clock, page-locked caller arrays and the staged default), and the device-resident
kernel time of constraint_eval_kernel and hyperplane_kernel (CUDA events)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
from problems import make_problem

n_obs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
q0, qd0, qdd0, _, obs = make_problem(100000, n_obs)
out = {"n_obs": n_obs}
for pinned in (True, False):
    p = ab.Planner(T=128, max_obstacles=max(n_obs, 1), device=0, pin_user_buffers=pinned)
    p.build(q0, qd0, qdd0, obs)
    m = p.m
    g, J = np.zeros(m), np.zeros(m * 7)
    rng = np.random.default_rng(1234)
    for timing in (True, False):
        p.set_kernel_timing(timing)
        wall, kern, inlib = [], [], []
        for i in range(220):
            x = rng.uniform(-1, 1, 7)
            t0 = time.perf_counter()
            p.eval_g_jac(x, g, J)
            wall.append((time.perf_counter() - t0) * 1e6)
            inlib.append(p.last_eval_host_us())
            if timing:
                kern.append(p.last_eval_ms() * 1e3)
        wall = wall[20:]
        key = ("pinned" if pinned else "staged") + ("_timed" if timing else "")
        out[key] = {"p50_us": float(np.percentile(wall, 50)), "p99_us": float(np.percentile(wall, 99)), "min_us": float(np.min(wall)),
                    "in_library_p50_us": float(np.percentile(inlib[20:], 50))}
        if timing:
            out[key]["kernel_p50_us"] = float(np.percentile(kern[20:], 50))
    if pinned:
        res = []
        for i in range(50):
            p.upload_x(rng.uniform(-1, 1, 7))
            p.eval_resident(None)
            res.append(p.last_eval_ms() * 1e3)
        out["resident_kernel_us"] = {"p50": float(np.percentile(res[5:], 50)), "min": float(np.min(res))}
        hy = []
        for i in range(10):
            p.build(q0, qd0, qdd0, obs)
            hy.append(p.last_build_ms()[2] * 1e3)
        out["hyperplane_kernel_us"] = {"p50": float(np.percentile(hy[2:], 50)), "min": float(np.min(hy))}
    p.close()
print("EVAL_LATENCY " + json.dumps(out))
