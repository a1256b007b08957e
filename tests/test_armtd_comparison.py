"""Second trajectory family: the ARMTD comparison planner — constant-acceleration
trajectory from offline JRS tables, forward occupancy only (kinova_planner_realtime_armtd_comparison, "KPA").
CPU: the oracle's restatement is pinned by properties.  GPU: the device path against the oracle."""
import numpy as np
import pytest
from scipy.optimize import linprog

import _oracle
import numeric_model as nm
from problems import DEBUG_K, armtd_displacement, make_jrs_tables, make_problem

T = 20           # the reference uses 100 (KPA/Parameters.h:17); a coarser grid keeps the CPU suite short
K_RANGE = np.array([np.pi / 24] * 7)


def problem(seed, n_obs):
    q0, qd0, _, q_des, obs = make_problem(seed, n_obs)
    return q0, qd0, q_des, obs, make_jrs_tables(qd0, K_RANGE, T)


@pytest.fixture(scope="module")
def built():
    q0, qd0, q_des, obs, jrs = problem(3, 5)
    o = _oracle.Oracle(T=T)
    o.build_armtd(q0, qd0, jrs, K_RANGE, obs)
    return o, q0, qd0, q_des, obs, jrs


def test_layout_and_bounds(built):
    o = built[0]
    n, m, nnz, _ = o.get_nlp_info()
    assert (n, m, nnz) == (7, 7 * T * 5 + 28, 7 * (7 * T * 5 + 28))        # KPA/NLPclass.cu:42-43
    xl, xu, gl, gu = o.get_bounds_info()
    assert np.all(gl[: 7 * T * 5] == -1e19) and np.all(gu[: 7 * T * 5] == 0)
    assert np.allclose(gu[-7:], np.array([1.3963] * 4 + [1.2218] * 3) - 2 * np.sqrt(2e-2 / 5.095620491878957))


def test_numeric_forward_kinematics_lies_in_link_reach_sets(built):
    o, q0, qd0, _, _, _ = built
    rng = np.random.default_rng(0)
    gens = o.link_generators()
    for trial in range(4):
        k = rng.uniform(-1, 1, 7)
        o.eval_g(k)
        centers = o.link_sliced_center()
        for s in rng.choice(T, 4, replace=False):
            t = (s + rng.uniform()) / T
            q = q0 + np.array([armtd_displacement(qd0[i], k[i] * K_RANGE[i], t) for i in range(7)])
            frames = nm.forward_kinematics(q)
            for link in range(7):
                R, p = frames[link]
                for corner in ((1, 1, 1), (-1, -1, 1), (1, -1, -1), (0, 0, 0)):
                    pt = p + R @ (nm.LINK_C[link] + np.array(corner) * nm.LINK_G[link])
                    G = gens[s, link]
                    res = linprog(np.zeros(6), A_eq=G, b_eq=pt - centers[s, link], bounds=[(-1 - 1e-9, 1 + 1e-9)] * 6, method="highs")
                    assert res.status == 0, (trial, s, link, corner)


def test_state_extrema_against_dense_sampling(built):
    o, q0, qd0, _, _, _ = built
    x = np.array([0.3, -0.8, 0.9, -0.2, 0.6, -0.5, 0.1])
    lim = o.eval_g(x)[-28:]
    ts = np.linspace(0, 1, 20001)
    for i in range(7):
        ka = x[i] * K_RANGE[i]
        q = q0[i] + armtd_displacement(qd0[i], ka, ts)
        v = np.gradient(q, ts)
        assert abs(lim[i] - q.min()) < 1e-6 and abs(lim[7 + i] - q.max()) < 1e-6
        assert abs(lim[14 + i] - v.min()) < 2e-3 and abs(lim[21 + i] - v.max()) < 2e-3


def test_jacobian(built):
    """Obstacle rows match finite differences.  The state rows are the reference's closed-form derivatives with respect
    to k_actual = k_range * k (KPA/Trajectory.cu:231-411 has no k_range factor): finite differences in k are k_range
    times them."""
    o = built[0]
    x = np.array([0.31, -0.42, 0.55, -0.2, 0.63, -0.5, 0.12])
    J, g0 = o.eval_jac_g(x), o.eval_g(x)
    for k in range(7):
        h = 1e-7
        xp, xm = x.copy(), x.copy()
        xp[k] += h
        xm[k] -= h
        gp, gm = o.eval_g(xp), o.eval_g(xm)
        fd = (gp - gm) / (2 * h)
        smooth = np.abs((gp - g0) / h - (g0 - gm) / h) < 1e-4
        obs_rows = np.arange(len(fd)) < len(fd) - 28
        assert np.abs(fd - J[:, k])[smooth & obs_rows].max() < 1e-5
        lim_rows = ~obs_rows & smooth
        assert np.abs(fd[lim_rows] - K_RANGE[k] * J[lim_rows, k]).max() < 1e-5
    off_diag = J[-28:].reshape(4, 7, 7) * (1 - np.eye(7))
    assert np.all(off_diag == 0)


def test_objective(built):
    o, q0, qd0, q_des, _, _ = built
    x = np.array([0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7])
    qp = q0 + qd0 * 0.5 + K_RANGE * x * 0.125                       # KPA/NLPclass.cu:197
    assert abs(o.eval_f(q_des, 0.5, x) - 10 * np.sum((q_des - qp) ** 2)) < 1e-9
    gr = o.eval_grad_f(q_des, 0.5, x)
    assert np.allclose(gr, 10 * 2 * (qp - q_des) * K_RANGE * 0.125)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_obs,T_", [(3, 5, 20), (8, 10, 100)])
def test_device_matches_oracle(seed, n_obs, T_, gpu_lib):
    import armour_b200 as ab
    from test_gpu_parity import assert_pz_equal, CTOL, REL
    q0, qd0, _, q_des, obs = make_problem(seed, n_obs)
    jrs = make_jrs_tables(qd0, K_RANGE, T_)
    o = _oracle.Oracle(T=T_)
    o.build_armtd(q0, qd0, jrs, K_RANGE, obs)
    p = ab.Planner(T=T_)
    p.build_armtd(q0, qd0, jrs, K_RANGE, obs)
    assert o.get_nlp_info() == p.get_nlp_info()
    for name in ("cos_q", "sin_q", "R", "R_t", "links"):
        for s in range(T_):
            for i in range(7):
                assert_pz_equal(o.get_pz(name, i, s), p.get_pz(name, i, s), "%s[%d,%d]" % (name, i, s))
    lg_o, lg_g = o.link_generators(), p.link_generators()
    assert np.abs(lg_o - lg_g).max() <= REL
    for a, b in zip(o.hyperplanes(), p.hyperplanes()):
        assert np.abs(a - b).max() <= REL
    rng = np.random.default_rng(seed)
    for x in (np.zeros(7) + 1e-3, DEBUG_K, rng.uniform(-1, 1, 7)):
        go, Jo = o.eval_g(x), o.eval_jac_g(x)
        gg, Jg = p.eval_g_jac(x)
        assert np.abs(go - gg).max() <= CTOL * max(1.0, np.abs(go).max())
        assert np.abs(Jo - Jg).max() <= CTOL * max(1.0, np.abs(Jo).max())
        assert o.check_feasible(go) == p.check_feasible(gg)
    for a, b in zip(o.get_bounds_info(), p.get_bounds_info()):
        assert np.array_equal(a, b)
    assert abs(o.eval_f(q_des, 0.5, DEBUG_K) - p.eval_f(q_des, 0.5, DEBUG_K)) <= 1e-12
    assert np.abs(o.eval_grad_f(q_des, 0.5, DEBUG_K) - p.eval_grad_f(q_des, 0.5, DEBUG_K)).max() <= 1e-12
    # the handle goes back to the ARMOUR path afterwards
    q0b, qd0b, qddb, _, obsb = make_problem(seed + 1, 3)
    if T_ == 20:
        p.build(q0b, qd0b, qddb, obsb)
        ob = _oracle.Oracle(T=T_)
        ob.build(q0b, qd0b, qddb, obsb)
        assert p.get_nlp_info() == ob.get_nlp_info()
        assert np.abs(ob.eval_g(DEBUG_K) - p.eval_g(DEBUG_K)).max() <= CTOL * 60
