"""Seed-indexed synthetic planning problems: shared by tests and bench.py."""
import numpy as np

STATE_LB = np.array([-np.pi, -2.41, -np.pi, -2.66, -np.pi, -2.23, -np.pi])  # continuous joints drawn in +-pi
STATE_UB = -STATE_LB
SPEED = np.array([1.3963] * 4 + [1.2218] * 3)


def make_problem(seed, n_obs=10):
    """q0, qd0, qdd0, q_des, obstacles[n_obs*12] (centre + 3 axis-aligned generators, column-major 3x4)."""
    rng = np.random.default_rng(seed)
    q0 = rng.uniform(STATE_LB, STATE_UB)
    qd0 = rng.uniform(-0.5 * SPEED, 0.5 * SPEED)
    qdd0 = rng.uniform(-1.0, 1.0, 7)
    q_des = q0 + rng.uniform(-np.pi / 6, np.pi / 6, 7)
    obs = np.zeros((n_obs, 12))
    for i in range(n_obs):
        c = rng.uniform([-0.8, -0.8, 0.1], [0.8, 0.8, 1.2])
        s = rng.uniform(0.02, 0.4, 3)
        obs[i, 0:3] = c
        obs[i, 3], obs[i, 7], obs[i, 11] = s[0] / 2, s[1] / 2, s[2] / 2
    return q0, qd0, qdd0, q_des, obs.ravel()


# state used by the reference's own harness (KPR/debug_script.m:29-31) and slice point (KPR/PZ_tests.cu:198)
DEBUG_Q0 = np.array([-1.0, -1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
DEBUG_QD0 = np.array([1.0, 1.0, 1.0, -1.0, -1.0, -1.0, -1.0])
DEBUG_QDD0 = np.full(7, 2.0)
DEBUG_K = np.array([0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7])

# example input in the reference's main() (KPR/armour_main.cu:19-34)
EXAMPLE_Q0 = np.array([0.6543, -0.0876, -0.4837, -1.2278, -1.5735, -1.0720, 0])
EXAMPLE_QDES = np.array([0.6831, 0.009488, -0.2471, -0.9777, -1.414, -0.9958, 0])
EXAMPLE_OBS = np.array([
    [-0.28239, -0.33281, 0.88069, 0.069825, 0, 0, 0, 0.09508, 0, 0, 0, 0.016624],
    [-0.19033, 0.035391, 1.3032, 0.11024, 0, 0, 0, 0.025188, 0, 0, 0, 0.014342],
    [0.67593, -0.085841, 0.43572, 0.17408, 0, 0, 0, 0.07951, 0, 0, 0, 0.18012],
    [0.75382, 0.51895, 0.4731, 0.030969, 0, 0, 0, 0.22312, 0, 0, 0, 0.22981],
    [0.75382, 0.51895, 0.4731, 0.030969, 0, 0, 0, 0.22312, 0, 0, 0, 0.22981],
    [-0.28239, -0.33281, 0.88069, 0.069825, 0, 0, 0, 0.09508, 0, 0, 0, 0.016624],
    [-0.19033, 0.035391, 1.3032, 0.11024, 0, 0, 0, 0.025188, 0, 0, 0, 0.014342],
    [0.67593, -0.085841, 0.43572, 0.17408, 0, 0, 0, 0.07951, 0, 0, 0, 0.18012],
    [0.75382, 0.51895, 0.4731, 0.030969, 0, 0, 0, 0.22312, 0, 0, 0, 0.22981],
    [0.75382, 0.51895, 0.4731, 0.030969, 0, 0, 0, 0.22312, 0, 0, 0, 0.22981],
]).ravel()


def armtd_displacement(qd0, k_actual, t):
    """Joint displacement of the ARMTD comparison planner's move-then-brake trajectory (KPA/Trajectory.h:6-16,
    KPA/Trajectory.cu:84-94): constant acceleration k until t = 0.5, then constant deceleration to rest at t = 1."""
    t = np.asarray(t, dtype=float)
    move = qd0 * t + 0.5 * k_actual * t ** 2
    peak, vpeak = qd0 * 0.5 + 0.125 * k_actual, qd0 + 0.5 * k_actual
    ts = t - 0.5
    brake = peak + vpeak * ts + 0.5 * (-vpeak / 0.5) * ts ** 2
    return np.where(t <= 0.5, move, brake)


def make_jrs_tables(qd0, k_range, T):
    """Synthetic stand-ins for the reference's offline JRS tables (kinova_planner_realtime_armtd_comparison/offline_jrs,
    binary .mat files that cannot travel): per joint and interval a zonotope c + g*k +- r enclosing cos / sin of the
    displacement.  Returns jrs[6, 7, T] in the reference's order c_cos, g_cos, r_cos, c_sin, g_sin, r_sin."""
    jrs = np.zeros((6, 7, T))
    ks = np.linspace(-1, 1, 9)
    for i in range(7):
        for s in range(T):
            ts = np.linspace(s / T, (s + 1) / T, 7)
            d = np.array([[armtd_displacement(qd0[i], k * k_range[i], t) for t in ts] for k in ks])   # [k, t]
            for base, f in ((0, np.cos), (3, np.sin)):
                v = f(d)
                g = float(np.mean(v[-1] - v[0]) / 2)
                c = float(np.mean(v - g * ks[:, None]))
                r = float(np.abs(v - c - g * ks[:, None]).max())
                jrs[base, i, s], jrs[base + 1, i, s], jrs[base + 2, i, s] = c, g, r
    return jrs


def saved_worlds():
    """The reference's 100 saved random worlds (kinova_src/saved_worlds/random/scene_*.csv), packed by
    tests/golden/make_saved_worlds.py: list of (name, start q, goal q, obstacles[n_obs * 12])."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "worlds_saved_random.npz"))
    out, off = [], 0
    for name, q, g, n in zip(d["names"], d["start"], d["goal"], d["n_obs"]):
        out.append((str(name), q.copy(), g.copy(), d["obstacles"][off:off + n].ravel().copy()))
        off += n
    return out
