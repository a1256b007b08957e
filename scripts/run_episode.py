"""Config 5 (BASELINE.json): receding-horizon episode under the 0.5 s planning deadline — a re-creation in Python of
the MATLAB loop (simulator_armtd.m:142-347, uarmtd_planner.m:85-435; MATLAB is not available): each step re-plans from
the state at t = 0.5 of the previous plan with a straight-line waypoint (lookahead 0.1, kinova_run_100_worlds.m:57),
at most 50 steps.  The solve uses the stand-in solver (Ipopt is not installed), so k differs from the reference's;
the latency distribution of build + solve through the C ABI is the point."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
import numeric_model as nm
from problems import make_problem, STATE_LB, STATE_UB

ap = argparse.ArgumentParser()
ap.add_argument("--seed", type=int, default=7)
ap.add_argument("--n-obs", type=int, default=10)
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
K_RANGE = np.pi / 48
q, _, _, _, obs = make_problem(args.seed, args.n_obs)
rng = np.random.default_rng(args.seed + 1)
goal = np.clip(q + rng.uniform(-0.6, 0.6, 7), STATE_LB, STATE_UB)
qd, qdd = np.zeros(7), np.zeros(7)
p = ab.Planner(T=128, max_obstacles=max(args.n_obs, 1))
p.build(q, qd, qdd, obs)   # untimed warm-up: the first launch loads the kernel module (~0.5 s); a planner service pays it once, not per step
lat, fails, consecutive = [], 0, 0
prev = None   # (q0, qd0, qdd0, k) of the last accepted plan, for the braking segment
for step in range(args.steps):
    dist_goal = np.linalg.norm(goal - q)
    if dist_goal < 0.05:
        break
    waypoint = q + (goal - q) / dist_goal * min(0.1 * np.sqrt(7), dist_goal)
    t0 = time.perf_counter()
    p.build(q, qd, qdd, obs)
    k, feas, it, ev = p.standin_solve(waypoint, 0.5)
    lat.append((time.perf_counter() - t0) * 1e3)
    if feas:
        consecutive = 0
        prev = (q.copy(), qd.copy(), qdd.copy(), k.copy())
        q, qd, qdd = [np.array(v) for v in zip(*[nm.bezier(q[i], qd[i], qdd[i], k[i] * K_RANGE, 0.5) for i in range(7)])]
    else:   # failed plan: finish the previous plan's braking half (it ends at rest), like the reference's agent
        fails += 1; consecutive += 1
        if prev is not None:
            q0, qd0, qdd0, kp = prev
            q = np.array([nm.bezier(q0[i], qd0[i], qdd0[i], kp[i] * K_RANGE, 1.0)[0] for i in range(7)])
        qd, qdd = np.zeros(7), np.zeros(7)
        prev = None
        if consecutive > 4:   # simulator_armtd.m:187-198
            break
print(json.dumps({"config": "receding-horizon episode (Python re-creation), T=128, %d obstacles" % args.n_obs, "steps": len(lat), "failed_plans": fails,
                  "reached_goal": bool(np.linalg.norm(goal - q) < 0.05), "latency_ms": {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)), "max": float(max(lat))},
                  "deadline_ms": 500.0, "within_deadline": bool(max(lat) < 500.0), "solver": "stand-in (Ipopt not installed)",
                  "warm_up": "one untimed build before the loop (kernel module load)"}))
