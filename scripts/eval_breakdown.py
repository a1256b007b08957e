import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
import armour_b200 as ab
from problems import make_problem
p = ab.Planner(T=128, max_obstacles=20, pin_user_buffers=bool(int(os.environ.get('PIN', '0'))))
q0, qd0, qdd0, _, obs = make_problem(100000, 20)
p.build(q0, qd0, qdd0, obs)
m = p.m; g = np.zeros(m); J = np.zeros(m * 7)
rng = np.random.default_rng(0)
def bench(f, n=200):
    ts = []
    for _ in range(n):
        x = rng.uniform(-1, 1, 7)
        t0 = time.perf_counter(); f(x); ts.append((time.perf_counter() - t0) * 1e6)
    return np.percentile(ts, 50), np.percentile(ts, 99)
p.upload_x(np.zeros(7))
print("kernel only + sync (resident x): p50 %.1f us p99 %.1f" % bench(lambda x: p.eval_resident(None)))
print("H2D x + kernel + sync:           p50 %.1f us p99 %.1f" % bench(lambda x: p.eval_resident(x)))
print("full eval_g_jac (D2H + memcpy):  p50 %.1f us p99 %.1f" % bench(lambda x: p.eval_g_jac(x, g, J)))
print("kernel ms", p.last_eval_ms())
a = np.zeros(m * 8); b = np.zeros(m * 8)
t0 = time.perf_counter()
for _ in range(200): np.copyto(a, b)
print("host memcpy of 8m doubles: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
