"""Device times of the BASELINE.json configurations that are parity cases rather than bench lines: (build_ms, reach kernel,
hyperplane kernel) by CUDA events and the end-to-end wall time of armour_build + one eval_g/eval_jac_g through the C ABI."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
from problems import make_problem
for label, T, n_obs, unc in (("configs[0] T=128, 10 obstacles", 128, 10, 0.03), ("configs[1] T=128, 20 obstacles", 128, 20, 0.03),
                             ("configs[3] T=512, 5 % uncertainty, 20 obstacles", 512, 20, 0.05), ("T=128, 40 obstacles (MAX_OBSTACLE_NUM)", 128, 40, 0.03)):
    p = ab.Planner(T=T, max_obstacles=n_obs, mass_uncertainty=unc, inertia_uncertainty=unc, pin_user_buffers=True)
    dev, wall = [], []
    for s in range(14):
        q0, qd0, qdd0, _, obs = make_problem(100000 + s, n_obs)
        if s == 0:
            p.build(q0, qd0, qdd0, obs); g, J = np.zeros(p.m), np.zeros(p.m * 7)
        t0 = time.perf_counter()
        p.build(q0, qd0, qdd0, obs)
        p.eval_g_jac(np.random.default_rng(s).uniform(-1, 1, 7), g, J)
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(p.last_build_ms())
    dev = np.array(dev[2:]); wall = np.array(wall[2:])
    print("%-50s m = %6d  build %.3f ms (reach %.3f, half-spaces %.3f)  build + eval end to end %.3f ms (medians)" % (label, p.m, np.median(dev[:, 0]), np.median(dev[:, 1]), np.median(dev[:, 2]), np.median(wall)), flush=True)
    p.close()
