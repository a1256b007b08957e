"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA path through the C ABI against the CPU
oracle on the same seeded inputs.

Tolerances are the ones BASELINE.json's north_star states:
  * monomial keys and keep/drop (generator reduction) choices: bit-exact
  * reach-set centres, coefficients (generators) and interval radii: 1e-9 relative; device radii and
    interval enclosures must additionally contain the oracle's (>=, <= on the end points)
  * constraint values and Jacobians: 1e-8
"""
import os
import zlib

import numpy as np
import pytest

import _oracle
import armour_b200 as ab
from problems import DEBUG_K, DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, EXAMPLE_OBS, EXAMPLE_Q0, make_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REL = 1e-9      # reach sets
CTOL = 1e-8     # constraints, Jacobians
TABLE_NAMES = ("cos_q", "sin_q", "R", "R_t", "qd_des", "qda_des", "qdda_des", "links", "u_nom", "u_nom_int")


def assert_pz_equal(a, b, what):
    """a = oracle, b = device."""
    assert (a["rows"], a["cols"]) == (b["rows"], b["cols"]), what
    assert np.array_equal(a["keys"], b["keys"]), "%s: monomial keys / reduction choices differ" % what
    scale = max(1.0, float(np.abs(a["center"]).max()))
    assert np.abs(a["center"] - b["center"]).max() <= REL * scale, what
    if len(a["keys"]):
        cs = max(scale, float(np.abs(a["coeffs"]).max()))
        assert np.abs(a["coeffs"] - b["coeffs"]).max() <= REL * cs, what
    ia, ib = a["independent"], b["independent"]
    assert np.all(ib >= ia), "%s: device radius does not contain the oracle's" % what
    assert np.all(ib - ia <= REL * np.maximum(np.abs(ia), 1e-12)), what


def compare_build(o, p, T):
    for name in TABLE_NAMES:
        for s in range(T):
            for i in range(7):
                assert_pz_equal(o.get_pz(name, i, s), p.get_pz(name, i, s), "%s[%d,%d]" % (name, i, s))
    tr_o, tr_g = o.torque_radius(), p.torque_radius()
    assert np.all(tr_g >= tr_o) and np.all(tr_g - tr_o <= REL * tr_o)
    lg_o, lg_g = o.link_generators(), p.link_generators()
    assert np.abs(lg_o - lg_g).max() <= REL * max(1.0, np.abs(lg_o).max())
    d = np.arange(3)
    assert np.all(lg_g[:, :, d, 3 + d] >= lg_o[:, :, d, 3 + d])      # diagonal block holds the interval radii
    (co, so), (cg, sg) = o.taylor_remainders(), p.taylor_remainders()
    for ref, dev in ((co, cg), (so, sg)):   # Boost-interval enclosures: device must contain the oracle's
        assert np.all(dev[..., 0] <= ref[..., 0]) and np.all(dev[..., 1] >= ref[..., 1])
        assert np.abs(dev - ref).max() <= REL * max(1e-3, np.abs(ref).max())


def compare_eval(o, p, x):
    go, Jo = o.eval_g(x), o.eval_jac_g(x)
    gg, Jg = p.eval_g_jac(x)
    assert np.abs(go - gg).max() <= CTOL * max(1.0, np.abs(go).max())
    assert np.abs(Jo - Jg).max() <= CTOL * max(1.0, np.abs(Jo).max())
    assert np.abs(o.link_sliced_center() - p.link_sliced_center()).max() <= REL
    # separate entry points serve the same numbers
    assert np.array_equal(p.eval_g(x), gg) and np.array_equal(p.eval_jac_g(x), Jg)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_config1_T128_10_obstacles(seed, gpu_lib):
    q0, qd0, qdd0, q_des, obs = make_problem(seed, 10)
    o = _oracle.Oracle(T=128)
    o.build(q0, qd0, qdd0, obs)
    p = ab.Planner(T=128)
    p.build(q0, qd0, qdd0, obs)
    assert o.get_nlp_info() == p.get_nlp_info() == (7, 9884, 69188, 0)
    compare_build(o, p, 128)
    Ao, do_, dlo = o.hyperplanes()
    Ag, dg, dlg = p.hyperplanes()
    assert max(np.abs(Ao - Ag).max(), np.abs(do_ - dg).max(), np.abs(dlo - dlg).max()) <= REL
    rng = np.random.default_rng(1234 + seed)
    for x in (np.zeros(7), DEBUG_K, rng.uniform(-1, 1, 7), rng.uniform(-1, 1, 7)):
        compare_eval(o, p, x)
    for a, b in zip(o.get_bounds_info(), p.get_bounds_info()):
        assert np.abs(a - b).max() <= REL * 100
    assert np.array_equal(o.get_starting_point(), p.get_starting_point())
    assert abs(o.eval_f(q_des, 0.5, DEBUG_K) - p.eval_f(q_des, 0.5, DEBUG_K)) <= 1e-12
    assert np.abs(o.eval_grad_f(q_des, 0.5, DEBUG_K) - p.eval_grad_f(q_des, 0.5, DEBUG_K)).max() <= 1e-12
    ir_o, jc_o = o.jac_structure()
    ir_g, jc_g = p.jac_structure()
    assert np.array_equal(ir_o, ir_g) and np.array_equal(jc_o, jc_g)
    g = p.eval_g(np.zeros(7))
    assert o.check_feasible(g) == p.check_feasible(g)


def test_config2_20_obstacles_reference_harness_state(gpu_lib):
    """State of KPR/debug_script.m:29-31, slice point of KPR/PZ_tests.cu:198, 20 obstacles (m = 18844)."""
    _, _, _, _, obs = make_problem(77, 20)
    o = _oracle.Oracle(T=128)
    o.build(DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, obs)
    p = ab.Planner(T=128)
    p.build(DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, obs)
    assert p.get_nlp_info() == (7, 18844, 131908, 0)
    compare_build(o, p, 128)
    rng = np.random.default_rng(1234)
    for x in [DEBUG_K] + [rng.uniform(-1, 1, 7) for _ in range(4)]:
        compare_eval(o, p, x)


def test_reference_example_input(gpu_lib):
    """The commented example in KPR/armour_main.cu:19-34: zero initial velocity and acceleration, which makes the
    k-independent stationary points 0/0 = NaN (KPR/Trajectory.cu:36-58) — comparisons with NaN must stay false."""
    z = np.zeros(7)
    o = _oracle.Oracle(T=128)
    o.build(EXAMPLE_Q0, z, z, EXAMPLE_OBS)
    p = ab.Planner(T=128)
    p.build(EXAMPLE_Q0, z, z, EXAMPLE_OBS)
    compare_build(o, p, 128)
    for x in (np.zeros(7), DEBUG_K, -DEBUG_K):
        compare_eval(o, p, x)


def test_config4_T512_five_percent_uncertainty(gpu_lib):
    q0, qd0, qdd0, _, obs = make_problem(5, 4)
    kw = dict(T=512, mass_uncertainty=0.05, inertia_uncertainty=0.05)
    o = _oracle.Oracle(**kw)
    o.build(q0, qd0, qdd0, obs)
    p = ab.Planner(**kw)
    p.build(q0, qd0, qdd0, obs)
    compare_build(o, p, 512)
    compare_eval(o, p, DEBUG_K)


def test_wide_k_range_grows_capacities(gpu_lib):
    """k_range = pi/24 (KPR/debug_script.m:35) produces lists beyond the default capacities: the build must grow
    them and still match, not fail or truncate."""
    kr = [np.pi / 24] * 7
    T = 128
    o = _oracle.Oracle(T=T, k_range=kr)
    o.build(DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, [])
    assert o.op_stats()["max_simplify_in"] > 4096 and o.op_stats()["max_simplify_out"] > 1024
    p = ab.Planner(T=T, k_range=kr, max_monomials=512, max_entries=2048)
    p.build(DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, [])
    compare_build(o, p, T)
    compare_eval(o, p, DEBUG_K)


def test_no_obstacles_and_max_obstacles(gpu_lib):
    q0, qd0, qdd0, _, obs40 = make_problem(9, 40)
    T = 8
    for obs in ([], obs40):
        o = _oracle.Oracle(T=T)
        o.build(q0, qd0, qdd0, obs)
        p = ab.Planner(T=T)
        p.build(q0, qd0, qdd0, obs)
        assert o.get_nlp_info() == p.get_nlp_info()
        compare_eval(o, p, DEBUG_K)
    p = ab.Planner(T=T)
    with pytest.raises(ab.ArmourError) as e:   # > MAX_OBSTACLE_NUM is rejected like KPR/armour_main.cu:66-72
        p.build(q0, qd0, qdd0, np.zeros(41 * 12))
    assert e.value.code == -1


def test_call_order_errors(gpu_lib):
    p = ab.Planner(T=8)
    with pytest.raises(ab.ArmourError) as e:
        p.eval_g(np.zeros(7))
    assert e.value.code == -4
    with pytest.raises(ab.ArmourError):
        ab.Planner(T=7)   # NUM_TIME_STEPS must be even (KPR/Parameters.h:16)


def test_batch_equals_single_builds(gpu_lib):
    """Independent problems in one launch give the same tables as one-at-a-time builds: monomial keys, coefficients and
    centres bit-identical.  Interval radii are block-wide round-up sums whose association follows the CTA shape (a sweep
    runs 128-thread CTAs, one plan two groups of 256), so they and the rows that contain them agree to 1e-12."""
    probs = [make_problem(s, 6) for s in (11, 12, 13)]
    T = 32
    pb = ab.Planner(T=T, batch=3)
    pb.build_batch(np.concatenate([q[0] for q in probs]), np.concatenate([q[1] for q in probs]), np.concatenate([q[2] for q in probs]),
                   np.concatenate([q[4] for q in probs]), 6)
    x = DEBUG_K
    for i, (q0, qd0, qdd0, _, obs) in enumerate(probs):
        ps = ab.Planner(T=T)
        ps.build(q0, qd0, qdd0, obs)
        pb.select_problem(i)
        g1, J1 = ps.eval_g_jac(x)
        g2, J2 = pb.eval_g_jac(x)
        np.testing.assert_allclose(g2, g1, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(J2, J1, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(pb.torque_radius(), ps.torque_radius(), rtol=1e-12, atol=0)
        for s in (0, T // 2, T - 1):
            a, b = ps.get_pz("u_nom", 3, s), pb.get_pz("u_nom", 3, s)
            assert np.array_equal(a["keys"], b["keys"]) and np.array_equal(a["coeffs"], b["coeffs"])


@pytest.mark.parametrize("pin", [False, True])
def test_eval_batch_equals_single_evals(gpu_lib, pin):
    """armour_eval_batch (one launch, one decision vector per problem) against one armour_eval_g_jac call per problem:
    constraint values and Jacobian rows bit-identical, through device staging and through page-locked caller arrays; a
    sub-range of the batch addresses the same problems."""
    n_obs, T, B = 5, 16, 4
    probs = [make_problem(s, n_obs) for s in (31, 32, 33, 34)]
    pb = ab.Planner(T=T, batch=B, pin_user_buffers=pin)
    pb.build_batch(np.concatenate([q[0] for q in probs]), np.concatenate([q[1] for q in probs]), np.concatenate([q[2] for q in probs]),
                   np.concatenate([q[4] for q in probs]), n_obs)
    rng = np.random.default_rng(77)
    xs = rng.uniform(-1, 1, (B, 7))
    G, V = pb.eval_batch(xs, want_jac=True)
    g_only = pb.eval_batch(xs)
    assert G.shape == (B, pb.m) and V.shape == (B, pb.m * 7) and np.array_equal(g_only, G)
    sub = pb.eval_batch(xs[1:3], first=1)
    assert np.array_equal(sub, G[1:3])
    for i in range(B):
        pb.select_problem(i)
        g, J = pb.eval_g_jac(xs[i])
        assert np.array_equal(g, G[i]) and np.array_equal(J.ravel(), V[i]), i
    G2 = np.zeros_like(G)
    pb.eval_batch_resident(xs, g=G2)      # Jacobians stay on the device, only g crosses PCIe
    assert np.array_equal(G2, G)
    for i in range(B):
        assert np.array_equal(pb.get_batch_jacobian(i), V[i]), i
    with pytest.raises(ab.ArmourError):
        pb.get_batch_jacobian(B)
    with pytest.raises(ab.ArmourError):
        pb.eval_batch(xs, first=1)        # range runs past the batch
    assert pb.last_eval_batch_ms() > 0
    pb.close()


def test_fused_half_space_tables_equal_the_separate_kernel(gpu_lib, monkeypatch):
    """Stage D is its own launch (hyperplane_kernel); ARMOUR_TUNE_FUSE_PLANES=1 runs it in the tail of reach_build_kernel
    instead (opt-in, measured not faster).  Same per-plane code: A, d, delta bit-identical, for one plan (two thread groups)
    and for a batch."""
    q0, qd0, qdd0, _, obs = make_problem(52, 20)
    p0 = ab.Planner(T=32)
    p0.build(q0, qd0, qdd0, obs)
    probs = [make_problem(s, 7) for s in (61, 62, 63)]
    args = [np.concatenate([q[k] for q in probs]) for k in (0, 1, 2, 4)]
    b0 = ab.Planner(T=16, batch=3)
    b0.build_batch(*args, 7)
    monkeypatch.setenv("ARMOUR_TUNE_FUSE_PLANES", "1")
    p1 = ab.Planner(T=32)
    p1.build(q0, qd0, qdd0, obs)
    b1 = ab.Planner(T=16, batch=3)
    b1.build_batch(*args, 7)
    assert p0.kernel_launches() == p1.kernel_launches() + 1
    for x, y in zip(p1.hyperplanes(), p0.hyperplanes()):
        assert x.size > 0 and np.array_equal(x, y)
    for i in range(3):
        b0.select_problem(i); b1.select_problem(i)
        for x, y in zip(b1.hyperplanes(), b0.hyperplanes()):
            assert np.array_equal(x, y)
        assert np.array_equal(b1.eval_g(DEBUG_K), b0.eval_g(DEBUG_K))


def test_solver_scratch_is_not_left_page_locked(gpu_lib):
    """Under cfg.pin_user_buffers every array handed to a callback is page-locked.  A solver frees its arrays when the solve
    returns, so the adapter's finalize_solution releases the registrations (armtd_NLP.hpp), and armour_standin_solve runs on the
    handle's own page-locked workspace.  A registration left on freed heap memory made cudaHostRegister fail for whatever array
    the allocator placed there next: the 8-GPU sweep of bench.py found it."""
    q0, qd0, qdd0, q_des, obs = make_problem(3, 5)
    for pin in (True, False):
        p = ab.Planner(T=16, pin_user_buffers=pin)
        p.build(q0, qd0, qdd0, obs)
        g, J = np.zeros(p.m), np.zeros(p.m * 7)
        p.eval_g_jac(DEBUG_K, g, J)
        assert p.pinned_buffer_count() == (2 if pin else 0)
        p.standin_solve(q_des, 0.5)
        assert p.pinned_buffer_count() == 0
        g2, J2 = np.zeros(p.m), np.zeros(p.m * 7)
        p.eval_g_jac(DEBUG_K, g2, J2)          # caller arrays are registered again on next use
        assert p.pinned_buffer_count() == (2 if pin else 0)
        assert np.array_equal(g, g2) and np.array_equal(J, J2)
        p.close()


def test_build_is_deterministic(gpu_lib):
    q0, qd0, qdd0, _, obs = make_problem(21, 10)
    p = ab.Planner(T=128)
    p.build(q0, qd0, qdd0, obs)
    g1, J1 = p.eval_g_jac(DEBUG_K)
    tr1 = p.torque_radius().copy()
    p.build(q0, qd0, qdd0, obs)
    g2, J2 = p.eval_g_jac(DEBUG_K)
    assert np.array_equal(g1, g2) and np.array_equal(J1, J2) and np.array_equal(tr1, p.torque_radius())


# ---- primitive-level parity: PZsparse arithmetic on random operands -------------------------------------------
def random_pz(rng, rows, cols, n, nvars=6, one_bit=range(7, 28)):
    """Random monomials of degree 1 in up to nvars - 1 of the 42 variables.  one_bit: which of the 21 one-bit error symbols
    (variable indices 7..27) may appear — the two operands of a product draw them from disjoint halves, because a product that
    squares a one-bit symbol overflows its key field (the reference would silently carry into the neighbouring variable; the
    device reports ARMOUR_E_NUMERIC, see test_degree_overflow_is_reported)."""
    dim = rows * cols
    keys = set()
    allowed = np.array([v for v in range(42) if v < 7 or v >= 28 or v in one_bit])
    while len(keys) < n:
        k = 0
        for v in rng.choice(allowed, size=rng.integers(1, nvars), replace=False):
            shift = 2 * v if v < 7 else (14 + (v - 7)) if v < 28 else 35 + 2 * (v - 28)
            k |= 1 << int(shift)
        keys.add(k)
    keys = np.array(sorted(keys), dtype=np.uint64)
    mag = 10.0 ** rng.uniform(-5, 0.5, size=(n, 1))     # straddles the 5e-4 threshold
    coeffs = rng.standard_normal((n, dim)) * mag
    coeffs[rng.random((n, dim)) < 0.25] = 0.0            # structural zeros as produced by element extraction
    return dict(rows=rows, cols=cols, keys=keys, coeffs=coeffs, center=rng.standard_normal(dim), independent=np.abs(rng.standard_normal(dim)) * 1e-2)


@pytest.mark.parametrize("op,shape_a,shape_b,na,nb", [
    ("mul", (3, 3), (3, 1), 3, 200), ("mul", (3, 3), (3, 1), 40, 60), ("mul", (3, 3), (3, 3), 30, 3), ("mul", (1, 1), (1, 1), 90, 20),
    ("mul", (3, 3), (3, 1), 0, 50), ("mul", (3, 3), (3, 1), 5, 0), ("mul", (1, 1), (1, 1), 0, 0),
    ("add", (3, 1), (3, 1), 150, 70), ("sub", (3, 1), (3, 1), 10, 300), ("add", (1, 1), (1, 1), 0, 12), ("sub", (1, 1), (1, 1), 33, 33),
    ("cross", (3, 1), (3, 1), 60, 25), ("cross", (3, 1), (3, 1), 2, 400), ("cross", (3, 1), (3, 1), 0, 7),
])
def test_pz_primitives(op, shape_a, shape_b, na, nb, gpu_lib):
    rng = np.random.default_rng(zlib.crc32(repr((op, shape_a, shape_b, na, nb)).encode()))
    p = ab.Planner(T=2)
    for trial in range(3):
        prod = op in ("mul", "cross")
        a = random_pz(rng, shape_a[0], shape_a[1], na, one_bit=range(7, 17) if prod else range(7, 28))
        b = random_pz(rng, shape_b[0], shape_b[1], nb, one_bit=range(17, 28) if prod else range(7, 28))
        if op in ("add", "sub") and trial == 1 and na and nb:   # force shared keys so that merging and cancellation happen
            m = min(na, nb) // 2
            b["keys"][:m] = a["keys"][:m]
            order = np.argsort(b["keys"], kind="stable")
            uniq = np.concatenate([[True], np.diff(b["keys"][order]) != 0])
            order = order[uniq]
            b["keys"], b["coeffs"] = b["keys"][order], b["coeffs"][order]
        ref = _oracle.pz_binary(op, a, b)
        dev = p.pz_binary(op, a, b)
        assert_pz_equal(ref, dev, "%s trial %d" % (op, trial))


def test_pz_self_subtraction_cancels(gpu_lib):
    rng = np.random.default_rng(5)
    a = random_pz(rng, 3, 1, 100)
    p = ab.Planner(T=2)
    r = p.pz_binary("sub", a, a)
    assert len(r["keys"]) == 0 and np.all(r["center"] == 0)
    assert np.all(r["independent"] >= 2 * a["independent"])


def test_pinned_user_buffers_give_identical_results(gpu_lib):
    """cfg.pin_user_buffers: the kernel writes the caller's arrays directly; same numbers as the staged path."""
    q0, qd0, qdd0, _, obs = make_problem(41, 12)
    a = ab.Planner(T=32)
    b = ab.Planner(T=32, pin_user_buffers=True)
    a.build(q0, qd0, qdd0, obs)
    b.build(q0, qd0, qdd0, obs)
    g, J = np.zeros(a.m), np.zeros(a.m * 7)
    for x in (DEBUG_K, np.zeros(7), -DEBUG_K):
        ga, Ja = a.eval_g_jac(x)
        b.eval_g_jac(x, g, J)
        assert np.array_equal(ga, g) and np.array_equal(Ja.ravel(), J)
    assert b.L.armour_release_host_buffers(b.h) == 0
    b.eval_g_jac(DEBUG_K, g, J)   # re-registers transparently
    assert np.array_equal(a.eval_g(DEBUG_K), g)


# ---- guards (round 2) --------------------------------------------------------------------------------------------------
def _scalar_pz(keys, coeffs):
    keys = np.array(keys, dtype=np.uint64)
    return dict(rows=1, cols=1, keys=keys, coeffs=np.array(coeffs, dtype=float).reshape(-1, 1), center=np.array([0.5]), independent=np.array([0.0]))


def test_degree_overflow_is_reported(gpu_lib):
    """The reference adds monomial keys without checking for a carry (KPR/PZsparse.cu:938-940); a degree that outgrows its
    field would silently corrupt the neighbouring variable.  The device checks every key addition: k_0^2 * k_0^2 (degree 4 in a
    2-bit field) and qde_0 * qde_0 (degree 2 in a 1-bit field) must fail with ARMOUR_E_NUMERIC; k_0^2 * k_0 and a product that
    fills every field to its maximum must pass."""
    p = ab.Planner(T=2)
    k0sq, k0, qde0 = 2, 1, 1 << 14
    for ka, kb in ((k0sq, k0sq), (qde0, qde0), (3 << 12, 1 << 12), (1 << 34, 1 << 34), (2 << 61, 2 << 61), (3 << 47, 1 << 47)):
        with pytest.raises(ab.ArmourError) as e:
            p.pz_binary("mul", _scalar_pz([ka], [1.0]), _scalar_pz([kb], [1.0]))
        assert e.value.code == -5, (ka, kb, e.value.code)
    r = p.pz_binary("mul", _scalar_pz([k0sq], [1.0]), _scalar_pz([k0], [1.0]))
    assert list(r["keys"]) == [k0, k0sq, 3]
    full_a = sum(2 << (2 * j) for j in range(7)) | sum(2 << (35 + 2 * j) for j in range(7)) | sum(2 << (49 + 2 * j) for j in range(7))
    full_b = sum(1 << (2 * j) for j in range(7)) | sum(1 << b for b in range(14, 35)) | sum(1 << (35 + 2 * j) for j in range(7)) | sum(1 << (49 + 2 * j) for j in range(7))
    r = p.pz_binary("mul", _scalar_pz([full_a], [1.0]), _scalar_pz([full_b], [1.0]))
    assert int(r["keys"][-1]) == full_a + full_b == (1 << 63) - 1
    p.close()


def test_pz_binary_rejects_unsupported_shapes(gpu_lib):
    p = ab.Planner(T=2)
    rng = np.random.default_rng(3)
    for op, sa, sb in (("mul", (3, 1), (3, 1)), ("mul", (4, 4), (4, 1)), ("add", (3, 3), (3, 3)), ("cross", (1, 1), (3, 1)), ("mul", (1, 1), (3, 1))):
        with pytest.raises(ab.ArmourError) as e:
            p.pz_binary(op, random_pz(rng, sa[0], sa[1], 3), random_pz(rng, sb[0], sb[1], 3))
        assert e.value.code == -1
    p.close()


def test_more_than_64_obstacles_is_rejected_at_create(gpu_lib):
    with pytest.raises(ab.ArmourError) as e:
        ab.Planner(T=8, max_obstacles=65)
    assert e.value.code == -1
    p = ab.Planner(T=8, max_obstacles=64)      # the limit itself works end to end
    q0, qd0, qdd0, _, obs = make_problem(77, 64)
    p.build(q0, qd0, qdd0, obs)
    o = _oracle.Oracle(T=8)
    o.build(q0, qd0, qdd0, obs)
    g, J = p.eval_g_jac(DEBUG_K)
    assert np.abs(g - o.eval_g(DEBUG_K)).max() <= 1e-8 and np.abs(J - o.eval_jac_g(DEBUG_K)).max() <= 1e-8
    p.close()


def test_k_only_table_capacity_grows_and_retries(gpu_lib):
    """UCAP / LCAP (k-only monomials kept per torque / link PZ) are runtime capacities: a build that overflows them doubles
    them and re-runs instead of failing.  Forced here by starting from 2-entry tables in a child process."""
    import os, subprocess, sys
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import numpy as np, armour_b200 as ab\nfrom problems import make_problem, DEBUG_K\n"
            "q0, qd0, qdd0, _, obs = make_problem(21, 4)\n"
            "p = ab.Planner(T=16); p.build(q0, qd0, qdd0, obs)\n"
            "g, J = p.eval_g_jac(DEBUG_K); print('SUM %%.17g %%.17g %%d' %% (g.sum(), np.abs(J).sum(), max(len(p.get_pz('u_nom', j, 8)['keys']) for j in range(7))))\n"
            % (os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")))
    outs = []
    for env in (dict(os.environ), dict(os.environ, ARMOUR_TUNE_UCAP="2", ARMOUR_TUNE_LCAP="2")):
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-800:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("SUM")][0])
    assert outs[0] == outs[1], outs
    assert int(outs[0].split()[-1]) > 2


def test_pinned_buffers_default_arrays_are_owned_by_the_planner(gpu_lib):
    """with pin_user_buffers the caller's arrays stay page-locked; temporaries allocated per call would be freed
    while registered and their addresses reused.  The wrapper now owns persistent buffers for the default-array path: many
    calls with interleaved allocations keep returning the right numbers."""
    q0, qd0, qdd0, _, obs = make_problem(43, 9)
    a = ab.Planner(T=16)
    b = ab.Planner(T=16, pin_user_buffers=True)
    a.build(q0, qd0, qdd0, obs)
    b.build(q0, qd0, qdd0, obs)
    rng = np.random.default_rng(2)
    junk = []
    for it in range(30):
        x = rng.uniform(-1, 1, 7)
        ga, Ja = a.eval_g_jac(x)
        gb, Jb = b.eval_g_jac(x)
        assert np.array_equal(ga, gb) and np.array_equal(Ja, Jb), it
        junk.append(np.zeros(rng.integers(1000, 200000)))      # churn the allocator between calls
        if it % 3 == 0:
            junk = junk[-2:]
    # more than 8 distinct caller arrays: entries are evicted, never the one resolved in the same call
    for it in range(12):
        g, J = np.zeros(b.m), np.zeros(b.m * 7)
        b.eval_g_jac(DEBUG_K, g, J)
        assert np.array_equal(g, a.eval_g(DEBUG_K)) and np.array_equal(J.reshape(-1, 7), a.eval_jac_g(DEBUG_K))
        assert b.L.armour_release_host_buffers(b.h) == 0 if it % 5 == 4 else True
        junk.append((g, J))        # keep them alive: the contract for caller-provided arrays
    a.close(); b.close()


# ---- the rest of the PZsparse public surface that runs on the device (round 2) ------------------------------------------
def _shuffled_with_repeats(rng, z, repeats):
    """an un-simplified monomial list: z's monomials in random order with `repeats` of them split into two parts"""
    keys, co = list(z["keys"]), [c.copy() for c in z["coeffs"]]
    for i in rng.choice(len(keys), size=min(repeats, len(keys)), replace=False):
        part = co[i] * rng.uniform(0.2, 0.8)
        co[i] = co[i] - part
        keys.append(keys[i]); co.append(part)
    order = rng.permutation(len(keys))
    return dict(z, keys=np.array(keys, dtype=np.uint64)[order], coeffs=np.array(co)[order])


@pytest.mark.parametrize("shape,n", [((1, 1), 40), ((3, 1), 300), ((3, 3), 25), ((3, 1), 0), ((3, 1), 2500)])
def test_pz_simplify_primitive(shape, n, gpu_lib):
    """PZsparse::simplify (KPR/PZsparse.cu:284-350) of an arbitrary list: unsorted, repeated keys, coefficients straddling
    the threshold.  The 2500-monomial case sorts in global memory (more candidates than the shared sort buffers hold)."""
    rng = np.random.default_rng(zlib.crc32(repr((shape, n)).encode()))
    p = ab.Planner(T=2)
    dummy = dict(rows=1, cols=1, keys=np.zeros(0, dtype=np.uint64), coeffs=np.zeros((0, 1)), center=np.zeros(1), independent=np.zeros(1))
    for trial in range(2):
        z = _shuffled_with_repeats(rng, random_pz(rng, shape[0], shape[1], n), n // 3)
        ref = _oracle.pz_binary("simplify", z, None)
        dev = p.pz_binary("simplify", z, dummy)
        assert_pz_equal(ref, dev, "simplify %s n=%d trial %d" % (shape, n, trial))
    p.close()


@pytest.mark.parametrize("row", [0, 1, 2])
def test_pz_add_one_dim_primitive(row, gpu_lib):
    rng = np.random.default_rng(40 + row)
    p = ab.Planner(T=2)
    for na, nb in ((120, 30), (0, 9), (50, 0)):
        a, b = random_pz(rng, 3, 1, na), random_pz(rng, 1, 1, nb)
        if na and nb:
            b["keys"][: nb // 2] = a["keys"][: nb // 2]        # shared keys: the scalar's terms merge into row `row`
            order = np.argsort(b["keys"], kind="stable")
            order = order[np.concatenate([[True], np.diff(b["keys"][order]) != 0])]
            b["keys"], b["coeffs"] = b["keys"][order], b["coeffs"][order]
        op = "add_one_dim%d" % row
        assert_pz_equal(_oracle.pz_binary(op, a, b), p.pz_binary(op, a, b), "%s na=%d nb=%d" % (op, na, nb))
    p.close()


@pytest.mark.parametrize("op", ["cross_const_first", "cross_const_second"])
def test_pz_cross_with_constant_primitive(op, gpu_lib):
    """cross(Eigen vector, PZ) and cross(PZ, Eigen vector) (KPR/PZsparse.cu:1118-1132, 1153-1167)"""
    rng = np.random.default_rng(zlib.crc32(op.encode()))
    p = ab.Planner(T=2)
    for n in (0, 7, 350):
        a = random_pz(rng, 3, 1, n)
        c = dict(rows=3, cols=1, keys=np.zeros(0, dtype=np.uint64), coeffs=np.zeros((0, 3)), center=rng.standard_normal(3) * [1.0, 1e-3, 0.2], independent=np.zeros(3))
        assert_pz_equal(_oracle.pz_binary(op, a, c), p.pz_binary(op, a, c), "%s n=%d" % (op, n))
    p.close()


def test_arena_guard_words_stay_intact(gpu_lib):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer_refusal.txt), so memory safety of the per-CTA arena
    is checked by a debug build of the library (`make canary`): a guard word behind every PZ slot, verified by the kernel after
    its last interval.  Runs a single plan (two thread groups) and a batch (sweep shape) with deliberately small monomial
    capacities, so that the grow-and-retry path — the place where an off-by-one would write past a slot — is exercised too."""
    import subprocess, sys
    lib = os.path.join(ROOT, "armour-dev_b200", "libarmour_b200_canary.so")
    assert os.path.exists(lib), "run __graft_entry__.build()"
    code = ("import sys, ctypes as C; sys.path[:0] = [%r, %r]\n"
            "import numpy as np, armour_b200 as ab\nab.LIB_PATH = %r\nfrom problems import make_problem, DEBUG_K\n"
            "def verified(p):\n    v = C.c_int(); assert p.L.armour_debug_canaries_verified(p.h, C.byref(v)) == 0; return v.value\n"
            "q0, qd0, qdd0, _, obs = make_problem(21, 4)\n"
            "for mono in (0, 96):\n"
            "    p = ab.Planner(T=16, max_monomials=mono); p.build(q0, qd0, qdd0, obs); g, J = p.eval_g_jac(DEBUG_K)\n"
            "    print('SINGLE %%d %%.17g' %% (verified(p), g.sum())); p.close()\n"
            "bp = [make_problem(70 + b, 4) for b in range(3)]\n"
            "pb = ab.Planner(T=16, batch=3, max_monomials=96)\n"
            "pb.build_batch(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]), np.concatenate([q[4] for q in bp]), 4)\n"
            "print('BATCH %%d' %% verified(pb))\n" % (os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests"), lib))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr[-800:]
    single = [l.split() for l in r.stdout.splitlines() if l.startswith("SINGLE")]
    assert len(single) == 2 and all(int(s[1]) >= 16 * 100 for s in single), r.stdout      # > 100 guard words per CTA, 16 CTAs
    assert single[0][2] == single[1][2]                                                  # the retried build gives the same numbers
    assert int([l.split()[1] for l in r.stdout.splitlines() if l.startswith("BATCH")][0]) >= 100


@pytest.mark.parametrize("shape_a,shape_b,small_is_a,small_vars,n_large", [
    ((3, 3), (3, 1), True, (5, 33, 40), 200),      # R_i^T * state: k_5, cosqe_5, sinqe_5 against a vector without joint-5 variables
    ((3, 3), (3, 1), True, (2, 30), 1),            # two surviving monomials, one-monomial vector
    ((3, 3), (3, 3), False, (6, 34, 41), 60),      # FK_R * R_i
    ((3, 3), (3, 1), False, (7, 14, 21), 55),      # FK_R * link box: generator symbols qde_0, qdae_0, qddae_0
    ((1, 1), (1, 1), True, (3,), 300),
    ((3, 3), (3, 1), True, (0, 28, 35), 340),      # lowest variables of every group: the shifted runs interleave most
])
def test_structured_product_equals_the_oracle(shape_a, shape_b, small_is_a, small_vars, n_large, gpu_lib):
    """Products whose operands share no variable and whose small operand has <= 3 single-variable monomials take the sort-free
    path of the engine (pz_mul_structured: closed-form positions from group bounds); results must equal the reference
    algorithm's sort + merge + threshold exactly like every other product."""
    rng = np.random.default_rng(zlib.crc32(repr((shape_a, shape_b, small_is_a, small_vars, n_large)).encode()))
    p = ab.Planner(T=2)
    shift = lambda v: 2 * v if v < 7 else (14 + (v - 7)) if v < 28 else 35 + 2 * (v - 28)
    for trial in range(3):
        ss, ls = (shape_a, shape_b) if small_is_a else (shape_b, shape_a)
        dim = ss[0] * ss[1]
        keys = np.array(sorted(1 << shift(v) for v in small_vars), dtype=np.uint64)
        mag = 10.0 ** rng.uniform(-4, 0, size=(len(keys), 1))
        small = dict(rows=ss[0], cols=ss[1], keys=keys, coeffs=rng.standard_normal((len(keys), dim)) * mag, center=rng.standard_normal(dim),
                     independent=np.abs(rng.standard_normal(dim)) * 1e-2)
        large = random_pz(rng, ls[0], ls[1], n_large, one_bit=[v for v in range(7, 28) if v not in small_vars])
        bad = np.zeros(len(large["keys"]), dtype=bool)          # drop monomials that use the small operand's variables
        for v in small_vars:
            bad |= ((large["keys"] >> np.uint64(shift(v))) & np.uint64(3 if (v < 7 or v >= 28) else 1)) != 0
        large["keys"], large["coeffs"] = large["keys"][~bad], large["coeffs"][~bad]
        a, b = (small, large) if small_is_a else (large, small)
        assert_pz_equal(_oracle.pz_binary("mul", a, b), p.pz_binary("mul", a, b), "structured trial %d" % trial)
    p.close()
