"""Pins the CPU oracle (oracle/) — the reference ships no golden vectors for this path, so the
oracle is checked by the properties the reference's own harness checks by eye (KPR/debug_script.m:94-123) plus the
authors' derivative test (KPR/armour_main.cu:268-273), and by hand-computable cases of the PZ algebra and of
Boost-style directed-rounding intervals."""
import ctypes as C

import numpy as np
import pytest
from scipy.optimize import linprog

import _oracle
import numeric_model as nm
from problems import DEBUG_K, DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, EXAMPLE_OBS, EXAMPLE_Q0, make_problem

K_RANGE = np.pi / 48
T = 32   # coarser than the reference's 128 keeps the CPU suite short; intervals are wider, the properties identical


@pytest.fixture(scope="module")
def built():
    q0, qd0, qdd0, q_des, obs = make_problem(4, 6)
    o = _oracle.Oracle(T=T)
    o.build(q0, qd0, qdd0, obs)
    return o, q0, qd0, qdd0, q_des, obs


def test_indicative_values_of_the_survey():
    """un-sliced u_nom centres and link-7 centre at s=64 for the debug_script state (T=128, pi/48);
    the survey's throw-away emulation is only good to ~1e-3."""
    o = _oracle.Oracle(T=128)
    o.build(DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, [])
    u = np.array([o.get_pz("u_nom", i, 64)["center"][0] for i in range(7)])
    assert np.allclose(u, [-17.29, -7.31, -21.32, 16.13, 7.90, 9.14, 8.42], atol=6e-3)
    assert np.allclose(o.get_pz("links", 6, 64)["center"], [-0.178, -0.637, 0.633], atol=1e-3)
    st = o.op_stats()
    assert st["n_mul"] == 128 * (21 + 2 * 238)       # 21 FK + 238 RNEA products per interval and RNEA pass
    assert st["max_simplify_in"] > st["max_simplify_out"] > 0


def in_zonotope(point, center, G):
    """LP: exists beta in [-1,1]^6 with center + G beta = point."""
    res = linprog(np.zeros(G.shape[1]), A_eq=G, b_eq=point - center, bounds=[(-1 - 1e-9, 1 + 1e-9)] * G.shape[1], method="highs")
    return res.status == 0


def test_numeric_forward_kinematics_lies_in_sliced_link_reach_sets(built):
    o, q0, qd0, qdd0, _, _ = built
    rng = np.random.default_rng(0)
    gens = o.link_generators()
    for trial in range(6):
        k = rng.uniform(-1, 1, 7)
        o.eval_g(k)
        centers = o.link_sliced_center()
        for s in rng.choice(T, 4, replace=False):
            t = (s + rng.uniform()) / T
            q, _, _ = nm.bezier(q0, qd0, qdd0, k * K_RANGE, t)
            frames = nm.forward_kinematics(q)
            for link in range(7):
                R, p = frames[link]
                for corner in ((1, 1, 1), (-1, 1, -1), (1, -1, -1), (-1, -1, 1), (0, 0, 0)):
                    pt = p + R @ (nm.LINK_C[link] + np.array(corner) * nm.LINK_G[link])
                    assert in_zonotope(pt, centers[s, link], gens[s, link]), (trial, s, link, corner)


def test_numeric_inverse_dynamics_lies_in_sliced_torque_reach_sets(built):
    """rnea(q, qd, qd, qdd) + armature * qdd must lie within the sliced u_nom centre +- its radius (zero tracking error)."""
    o, q0, qd0, qdd0, _, _ = built
    rng = np.random.default_rng(1)
    for trial in range(6):
        k = rng.uniform(-1, 1, 7)
        g = o.eval_g(k)[: 7 * T].reshape(T, 7)
        for s in range(0, T, 3):
            radius = np.array([o.get_pz("u_nom", j, s)["independent"][0] for j in range(7)])
            t = (s + rng.uniform()) / T
            q, qd, qdd = nm.bezier(q0, qd0, qdd0, k * K_RANGE, t)
            u = nm.rnea(q, qd, qd, qdd)
            assert np.all(np.abs(u - g[s]) <= radius + 1e-9), (trial, s, u - g[s], radius)
    # the robust torque radius adds the disturbance bound on top of it
    tr = o.torque_radius()
    assert np.all(tr >= np.array([[o.get_pz("u_nom", j, s)["independent"][0] for j in range(7)] for s in range(T)]))


def test_interval_rnea_radius_dominates_nominal(built):
    o = built[0]
    for s in (0, T // 2, T - 1):
        for j in range(7):
            d = o.get_pz("u_nom_int", j, s)
            assert len(d["keys"]) == 0 and d["center"][0] == 0.0     # polynomial parts cancel exactly (KPR/armour_main.cu:135-137)
            assert d["independent"][0] > 0


def test_jacobian_matches_finite_differences(built):
    """The authors' derivative_test: perturbation 1e-8, tolerance 1e-6 (KPR/armour_main.cu:271-272)."""
    o = built[0]
    rng = np.random.default_rng(2)
    for trial in range(3):
        x = rng.uniform(-0.9, 0.9, 7)
        J = o.eval_jac_g(x)
        g0 = o.eval_g(x)
        for k in range(7):
            h = 1e-7
            xp, xm = x.copy(), x.copy()
            xp[k] += h
            xm[k] -= h
            gp, gm = o.eval_g(xp), o.eval_g(xm)
            fd = (gp - gm) / (2 * h)
            # rows that are a max over planes / extrema candidates have kinks: skip rows where the one-sided
            # differences disagree (the active piece switches inside the stencil)
            fwd, bwd = (gp - g0) / h, (g0 - gm) / h
            smooth = np.abs(fwd - bwd) < 1e-4
            # Reference quirk, reproduced on purpose: returnJointVelocityExtremumGradient reports d/dk = 1 * k_range
            # when the extremum sits at t = 1 (KPR/Trajectory.cu:507-509,524-526) although qd_des(1; k) == 0 for
            # every k.  Those velocity rows cannot match finite differences.
            quirk = np.zeros(len(fd), dtype=bool)
            quirk[-14:] = (np.abs(J[-14:, k]) == K_RANGE) & (np.abs(fd[-14:]) < 1e-9)
            smooth &= ~quirk
            assert np.abs(fd - J[:, k])[smooth].max() < 1e-5, (trial, k)
            assert smooth.mean() > 0.97


def test_limit_rows_against_dense_sampling(built):
    o, q0, qd0, qdd0, _, _ = built
    x = np.array([0.3, -0.8, 0.9, -0.2, 0.6, -0.5, 0.1])
    g = o.eval_g(x)
    lim = g[-28:]
    ts = np.linspace(0, 1, 20001)
    for i in range(7):
        q, qd, _ = nm.bezier(q0[i], qd0[i], qdd0[i], x[i] * K_RANGE, ts)
        assert abs(lim[i] - q.min()) < 1e-6 and abs(lim[7 + i] - q.max()) < 1e-6
        assert abs(lim[14 + i] - qd.min()) < 1e-6 and abs(lim[21 + i] - qd.max()) < 1e-6


def test_constraint_layout_and_bounds(built):
    o, q0, qd0, qdd0, q_des, obs = built
    n, m, nnz, nh = o.get_nlp_info()
    assert (n, m, nnz, nh) == (7, 7 * T + 7 * T * 6 + 28, 7 * (7 * T + 7 * T * 6 + 28), 0)    # KPR/NLPclass.cu:47-49,76
    xl, xu, gl, gu = o.get_bounds_info()
    assert np.all(xl == -1) and np.all(xu == 1)
    assert np.all(gl[7 * T: 7 * T + 7 * T * 6] == -1e19) and np.all(gu[7 * T: 7 * T + 7 * T * 6] == 0)
    tr = o.torque_radius()
    lim = np.array([56.7] * 4 + [29.4] * 3)
    assert np.allclose(gu[: 7 * T].reshape(T, 7), lim - tr) and np.allclose(gl[: 7 * T].reshape(T, 7), -lim + tr)
    ir, jc = o.jac_structure()
    assert np.array_equal(ir, np.repeat(np.arange(m), 7)) and np.array_equal(jc, np.tile(np.arange(7), m))
    # objective: 10 * sum wrap?(q_des - q(t_plan))^2 and its gradient (KPR/NLPclass.cu:207-267)
    x = np.array([0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7])
    f = o.eval_f(q_des, 0.5, x)
    gr = o.eval_grad_f(q_des, 0.5, x)
    for k in range(7):
        xp, xm = x.copy(), x.copy()
        xp[k] += 1e-6
        xm[k] -= 1e-6
        assert abs((o.eval_f(q_des, 0.5, xp) - o.eval_f(q_des, 0.5, xm)) / 2e-6 - gr[k]) < 1e-6
    qp = np.array([nm.bezier(q0[i], qd0[i], qdd0[i], x[i] * K_RANGE, 0.5)[0] for i in range(7)])
    assert abs(f - 10 * np.sum((q_des - qp) ** 2)) < 1e-9     # |q_des - q| < pi here, so no wrapping


def test_obstacle_rows_are_separating_plane_distances(built):
    """g_obs = -max over the 72 signed plane distances (KPR/CollisionChecking.cu:230-283): recompute from the tables."""
    o = built[0]
    x = DEBUG_K
    g = o.eval_g(x)
    A, d, delta = o.hyperplanes()
    c = o.link_sliced_center()
    n_obs = A.shape[2]
    gobs = g[7 * T: 7 * T + 7 * T * n_obs].reshape(7, T, n_obs)
    dot = np.einsum("tlopa,tla->tlop", A, c)
    valid = np.linalg.norm(A, axis=-1) > 0
    pos = np.where(valid, dot - (d + delta), -1e8)
    neg = np.where(valid, -dot - (-d + delta), -1e8)
    best = np.maximum(pos, neg).max(axis=-1)
    assert np.allclose(gobs, -best.transpose(1, 0, 2), atol=1e-12)
    # each valid plane normal is a unit vector orthogonal to its two generators' cross product direction
    assert np.allclose(np.linalg.norm(A, axis=-1)[valid], 1.0)


def test_zero_velocity_start_has_no_nan(built):
    o = _oracle.Oracle(T=8)
    z = np.zeros(7)
    o.build(EXAMPLE_Q0, z, z, EXAMPLE_OBS)     # KPR/armour_main.cu:19-34
    g = o.eval_g(np.zeros(7))
    assert np.all(np.isfinite(g)) and np.all(np.isfinite(o.eval_jac_g(np.zeros(7))))
    assert np.all(np.isfinite(o.torque_radius()))


# ---- PZ algebra -----------------------------------------------------------------------------------------------
def K(j):
    return np.uint64(1 << (2 * j))


def scalar(center, terms, ind=0.0):
    keys = np.array(sorted(terms), dtype=np.uint64)
    return dict(rows=1, cols=1, keys=keys, coeffs=np.array([[terms[int(k)]] for k in keys], dtype=float).reshape(len(keys), 1), center=[center], independent=[ind])


def test_scalar_product_by_hand():
    """(1 + 2 k0 + 3 k1) * (4 + 5 k0) = 4 + 13 k0 + 12 k1 + 10 k0^2 + 15 k0 k1 ; keys add (KPR/PZsparse.cu:938-940)."""
    a = scalar(1.0, {int(K(0)): 2.0, int(K(1)): 3.0})
    b = scalar(4.0, {int(K(0)): 5.0})
    r = _oracle.pz_binary("mul", a, b)
    got = dict(zip((int(k) for k in r["keys"]), r["coeffs"][:, 0]))
    assert r["center"][0] == 4.0
    assert got == {1: 13.0, 2: 10.0, 4: 12.0, 5: 15.0}
    assert list(r["keys"]) == sorted(r["keys"])


def test_threshold_moves_small_monomials_into_the_radius():
    a = scalar(0.0, {int(K(0)): 1e-2, int(K(1)): 1.0}, ind=0.5)
    b = scalar(0.0, {int(K(2)): 4e-2}, ind=0.25)
    r = _oracle.pz_binary("mul", a, b)
    # k0*k2 has coefficient 4e-4 <= 5e-4 -> dropped; k1*k2 = 4e-2 kept
    assert [int(k) for k in r["keys"]] == [int(K(1)) + int(K(2))]
    expect = 0.5 * 0.25 + (0 + 1e-2 + 1.0) * 0.25 + 0.5 * (0 + 4e-2) + 4e-4      # KPR/PZsparse.cu:944-989 + simplify
    assert abs(r["independent"][0] - expect) < 1e-15
    exactly = scalar(0.0, {int(K(0)): 5e-4})
    assert len(_oracle.pz_binary("simplify", exactly)["keys"]) == 0            # "<=" threshold (KPR/PZsparse.cu:309)


def test_sum_difference_and_cancellation():
    a = scalar(1.0, {int(K(0)): 2.0, int(K(3)): -1.0}, ind=0.1)
    b = scalar(-3.0, {int(K(0)): 2.0, int(K(5)): 7.0}, ind=0.2)
    s = _oracle.pz_binary("add", a, b)
    d = _oracle.pz_binary("sub", a, b)
    assert s["center"][0] == -2.0 and d["center"][0] == 4.0
    assert dict(zip((int(k) for k in s["keys"]), s["coeffs"][:, 0])) == {int(K(0)): 4.0, int(K(3)): -1.0, int(K(5)): 7.0}
    assert dict(zip((int(k) for k in d["keys"]), d["coeffs"][:, 0])) == {int(K(3)): -1.0, int(K(5)): -7.0}   # k0 cancels
    assert abs(s["independent"][0] - 0.3) < 1e-16 and abs(d["independent"][0] - 0.3) < 1e-16


def test_cross_product_matches_numpy_on_slices():
    rng = np.random.default_rng(3)

    def vec(n):
        keys = np.array(sorted(rng.choice(2 ** 14, n, replace=False)), dtype=np.uint64)
        keys = keys & np.uint64(0x1555)        # degree <= 1 per k_j so that products stay within the 2-bit fields
        keys = np.unique(keys[keys > 0])
        return dict(rows=3, cols=1, keys=keys, coeffs=rng.standard_normal((len(keys), 3)), center=rng.standard_normal(3), independent=np.zeros(3))

    a, b = vec(12), vec(9)
    r = _oracle.pz_binary("cross", a, b, threshold=0.0)

    def slice_at(z, x):
        v = np.array(z["center"], dtype=float)
        for k, c in zip(z["keys"], z["coeffs"]):
            mon = 1.0
            for j in range(7):
                mon *= x[j] ** ((int(k) >> (2 * j)) & 3)
            v = v + c * mon
        return v

    for _ in range(5):
        x = rng.uniform(-1, 1, 7)
        assert np.allclose(slice_at(r, x), np.cross(slice_at(a, x), slice_at(b, x)), atol=1e-12)


def test_key_layout():
    """KPR/PZsparse.h:23-40: k_j 2 bits at 2j; qde/qdae/qddae 1 bit at 14/21/28 + j; cosqe/sinqe 2 bits at 35/49 + 2j."""
    o = _oracle.Oracle(T=4)
    o.build(DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, [])
    for j in range(7):
        assert list(o.get_pz("qd_des", j, 1)["keys"]) == [1 << (2 * j), 1 << (14 + j)]
        assert list(o.get_pz("qda_des", j, 1)["keys"]) == [1 << (2 * j), 1 << (21 + j)]
        assert list(o.get_pz("qdda_des", j, 1)["keys"]) == [1 << (2 * j), 1 << (28 + j)]
        assert list(o.get_pz("cos_q", j, 1)["keys"]) == [1 << (2 * j), 1 << (35 + 2 * j)]
        assert list(o.get_pz("sin_q", j, 1)["keys"]) == [1 << (2 * j), 1 << (49 + 2 * j)]
        R = o.get_pz("R", j, 1)
        assert list(R["keys"]) == [1 << (2 * j), 1 << (35 + 2 * j), 1 << (49 + 2 * j)] and (R["rows"], R["cols"]) == (3, 3)
        Rt = o.get_pz("R_t", j, 1)
        assert np.allclose(R["center"].reshape(3, 3), Rt["center"].reshape(3, 3).T)


def test_joint_reach_sets_enclose_the_trajectory():
    """cos/sin(q_des(t;k)) and qd_des, qdd_des must lie inside the sliced PZs +- their error generators for every t in
    the interval (the property makePolyZono exists for, KPR/Trajectory.cu:63-254)."""
    T2 = 16
    o = _oracle.Oracle(T=T2)
    q0, qd0, qdd0, _, _ = make_problem(8, 0)
    o.build(q0, qd0, qdd0, [])
    rng = np.random.default_rng(4)
    for _ in range(20):
        j, s = rng.integers(7), rng.integers(T2)
        k = rng.uniform(-1, 1)
        t = (s + rng.uniform()) / T2
        q, qd, qdd = nm.bezier(q0[j], qd0[j], qdd0[j], k * K_RANGE, t)
        for name, val in (("cos_q", np.cos(q)), ("sin_q", np.sin(q)), ("qd_des", qd), ("qdda_des", qdd)):
            z = o.get_pz(name, j, s)
            kk = dict(zip((int(x) for x in z["keys"]), z["coeffs"][:, 0]))
            mid = z["center"][0] + kk.get(1 << (2 * j), 0.0) * k
            rad = z["independent"][0] + sum(abs(v) for key, v in kk.items() if key != (1 << (2 * j)))
            assert abs(val - mid) <= rad + 1e-12, (name, j, s)


# ---- directed-rounding intervals (exported remainders must be sound) ----------------------------------------------
def test_taylor_remainders_are_sound_enclosures():
    """cos(q) - cos(qc) + (q - qc) sin(qc)... the exported interval is the remainder  -d sin(qc) - 0.5 cos(xi) d^2 ;
    check with high-precision sampling that cos(qc + d) - cos(qc) lies inside it for admissible d."""
    import mpmath as mp
    mp.mp.dps = 40
    T2 = 8
    o = _oracle.Oracle(T=T2)
    q0, qd0, qdd0, _, _ = make_problem(15, 0)
    o.build(q0, qd0, qdd0, [])
    crem, srem = o.taylor_remainders()
    assert np.all(crem[..., 0] <= crem[..., 1]) and np.all(srem[..., 0] <= srem[..., 1])
    rng = np.random.default_rng(6)
    for _ in range(30):
        j, s = rng.integers(7), rng.integers(T2)
        k = rng.uniform(-1, 1)
        t = (s + rng.uniform()) / T2
        q, _, _ = nm.bezier(q0[j], qd0[j], qdd0[j], k * K_RANGE, t)
        z = o.get_pz("cos_q", j, s)
        kk = dict(zip((int(x) for x in z["keys"]), z["coeffs"][:, 0]))
        # value of the PZ without its error generator; the true cosine must be within the remainder's width of it
        mid = z["center"][0] + kk.get(1 << (2 * j), 0.0) * k
        width = (crem[j, s, 1] - crem[j, s, 0]) / 2
        assert abs(float(mp.cos(mp.mpf(float(q)))) - mid) <= width + 1e-15
