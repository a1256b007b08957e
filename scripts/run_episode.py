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
ap.add_argument("--worlds", action="store_true", help="run the reference's 100 saved random worlds instead of one synthetic episode")
args = ap.parse_args()
K_RANGE = np.pi / 48


def episode(p, q, goal, obs, max_steps):
    """One receding-horizon episode; returns (per-step latencies ms, failed plans, reached goal, steps)."""
    qd, qdd = np.zeros(7), np.zeros(7)
    lat, fails, consecutive = [], 0, 0
    prev = None   # (q0, qd0, qdd0, k) of the last accepted plan, for the braking segment
    for step in range(max_steps):
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
    return lat, fails, bool(np.linalg.norm(goal - q) < 0.05), float(np.linalg.norm(goal - q))


if args.worlds:
    # kinova_run_100_worlds.m re-creation: the reference's 100 saved random worlds (start, goal, 5-14 boxes), <= 50 plans each
    from problems import saved_worlds
    worlds = saved_worlds()
    p = ab.Planner(T=128, max_obstacles=max(len(w[3]) // 12 for w in worlds))
    p.build(worlds[0][1], np.zeros(7), np.zeros(7), worlds[0][3])   # untimed warm-up (kernel module load)
    all_lat, reached, stopped, total_fail, d0, d1 = [], 0, 0, 0, [], []
    for name, q0, goal, obs in worlds:
        lat, fails, ok, dist = episode(p, q0.copy(), goal, obs, args.steps)
        all_lat += lat; reached += ok; total_fail += fails; stopped += (not ok); d0.append(float(np.linalg.norm(goal - q0))); d1.append(dist)
    all_lat = np.array(all_lat)
    print(json.dumps({"config": "100 saved random worlds of the reference (kinova_src/saved_worlds/random), receding-horizon episodes (Python re-creation of kinova_run_100_worlds.m), T=128, 5-14 obstacles",
                      "worlds": len(worlds), "goals_reached": int(reached), "episodes_ended_by_the_%d-plan_cap_or_5_failed_plans" % args.steps: int(stopped),
                      "joint_space_distance_to_goal_rad": {"median_at_start": float(np.median(d0)), "median_at_end": float(np.median(d1))}, "plan_steps": int(all_lat.size), "failed_plans": int(total_fail),
                      "latency_ms": {"p50": float(np.percentile(all_lat, 50)), "p90": float(np.percentile(all_lat, 90)), "p99": float(np.percentile(all_lat, 99)), "max": float(all_lat.max())},
                      "deadline_ms": 500.0, "within_deadline": bool(all_lat.max() < 500.0), "solver": "stand-in (Ipopt not installed): its k and therefore the success counts differ from the reference's",
                      "warm_up": "one untimed build before the first world (kernel module load)"}))
    sys.exit(0)

q, _, _, _, obs = make_problem(args.seed, args.n_obs)
rng = np.random.default_rng(args.seed + 1)
goal = np.clip(q + rng.uniform(-0.6, 0.6, 7), STATE_LB, STATE_UB)
p = ab.Planner(T=128, max_obstacles=max(args.n_obs, 1))
p.build(q, np.zeros(7), np.zeros(7), obs)   # untimed warm-up: the first launch loads the kernel module (~0.5 s); a planner service pays it once, not per step
lat, fails, ok, _ = episode(p, q, goal, obs, args.steps)
print(json.dumps({"config": "receding-horizon episode (Python re-creation), T=128, %d obstacles" % args.n_obs, "steps": len(lat), "failed_plans": fails,
                  "reached_goal": ok, "latency_ms": {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)), "max": float(max(lat))},
                  "deadline_ms": 500.0, "within_deadline": bool(max(lat) < 500.0), "solver": "stand-in (Ipopt not installed)",
                  "warm_up": "one untimed build before the loop (kernel module load)"}))
