"""Robust low-level controller (KRC = kinova_robust_controllers_mex).

CPU part pins the oracle restatement (oracle/oracle_controller.cpp) by properties — the reference ships no recorded
controller outputs: nominal RNEA against an independent numeric Newton-Euler of the same robot, interval RNEA against
sampled models inside the uncertainty box, the robust input against its closed form.  GPU part compares the device path
(armour_controller_* C ABI) with the oracle on the same seeded inputs.

Tolerance (floating point): 1e-10 absolute + relative.  Both sides follow the same operation order with the same
rounding; the only difference is sin/cos of the joint angle (CUDA libm vs glibc, <= 2 ulp)."""
import ctypes
import os

import numpy as np
import pytest

import _oracle
import numeric_model as nm

HERE = os.path.dirname(os.path.abspath(__file__))
MODEL = os.path.join(HERE, "golden", "kinova_without_gripper.txt")
GOLDEN = os.path.join(HERE, "golden", "controller_64.npz")
TOL = 1e-10
# defaults of uarmtd_robust_CBF_LLC.m:6-13
KR, ALPHA, V_MAX, R_THR = 5.0, 10.0, 1e-2, 0.0
KP, KI, MAX_ERR = [28.1037, 2.0], [2.0, 0.2], 1e-5


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


def states(seed, count, err=0.05):
    """Measured state near a desired state, like a tracking controller sees."""
    rng = np.random.default_rng(seed)
    qd = rng.uniform(-np.pi, np.pi, (count, 7))
    qd_d = rng.uniform(-1.0, 1.0, (count, 7))
    qd_dd = rng.uniform(-2.0, 2.0, (count, 7))
    q = qd + rng.uniform(-err, err, (count, 7))
    q_d = qd_d + rng.uniform(-err, err, (count, 7))
    return q, q_d, qd, qd_d, qd_dd


# ------------------------------------------------------------------------------------------------- oracle (CPU)
def test_generated_model_file_matches_planner_constants():
    """Body-CoM conversion (KRC/robot_models.cpp:124-151) of the generated file gives back the planner's constants:
    inertia about the CoM, zero first moment, masses, armature."""
    o = _oracle.OracleController(MODEL)
    m = o.model()
    assert o.n == 7
    for i in range(7):
        assert abs(m[i, 6] - nm.MASS[i]) == 0
        assert np.abs(m[i, 7:16].reshape(3, 3) - nm.INERTIA[i]).max() < 1e-15
        assert np.abs(m[i, 16:25]).max() < 1e-15
        assert m[i, 25] == nm.ARMATURE[i]


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="reference tree not present")
def test_generated_model_file_equals_reference_file():
    ref = "/root/reference/kinova_src/kinova_simulator_interfaces/kinova_robust_controllers_mex/kinova_without_gripper.txt"
    a, b = _oracle.OracleController(MODEL).model(), _oracle.OracleController(ref).model()
    assert np.abs(a - b).max() < 1e-13


def test_nominal_rnea_equals_independent_newton_euler():
    o = _oracle.OracleController(MODEL)
    rng = np.random.default_rng(1)
    for _ in range(20):
        q, qd, qda, qdd = (rng.uniform(-1, 1, 7) * s for s in (3.0, 1.0, 1.0, 2.0))
        assert np.abs(o.rnea(q, qd, qda, qdd) - nm.rnea(q, qd, qda, qdd)).max() < 1e-12


def test_interval_rnea_encloses_sampled_models():
    eps = 0.03
    o = _oracle.OracleController(MODEL, eps)
    rng = np.random.default_rng(2)
    mass0, inertia0 = nm.MASS.copy(), nm.INERTIA.copy()
    try:
        for _ in range(5):
            q, qd, qda, qdd = (rng.uniform(-1, 1, 7) * s for s in (3.0, 1.0, 1.0, 2.0))
            t = o.rnea(q, qd, qda, qdd, interval=True)
            nominal = o.rnea(q, qd, qda, qdd)
            assert np.all(t[:, 0] <= nominal) and np.all(nominal <= t[:, 1])
            width = t[:, 1] - t[:, 0]
            assert np.all(width > 0) and np.all(width < 5.0)
            for _ in range(20):
                nm.MASS = mass0 * (1 + rng.uniform(-eps, eps, 7))
                nm.INERTIA = inertia0 * (1 + rng.uniform(-eps, eps, (7, 3, 3)))
                s = nm.rnea(q, qd, qda, qdd)
                assert np.all(t[:, 0] - 1e-12 <= s) and np.all(s <= t[:, 1] + 1e-12)
    finally:
        nm.MASS, nm.INERTIA = mass0, inertia0


def test_robust_input_closed_forms():
    o = _oracle.OracleController(MODEL)
    q, q_d, qd, qd_d, qd_dd = states(3, 32)
    u, un, v, ui, Vs, outside = o.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd)
    assert outside == 0
    assert np.array_equal(u, un - v)
    r = (qd_d - q_d) + KR * (qd - q)
    r_norm = np.linalg.norm(r, axis=1)
    bound = np.maximum(np.abs(ui[..., 0] - un), np.abs(ui[..., 1] - un))
    lam = np.maximum(0.0, -ALPHA * (V_MAX - Vs) / r_norm + np.linalg.norm(bound, axis=1))
    assert close(v, -lam[:, None] * r / r_norm[:, None])
    # V_sup encloses 0.5 r' M(q) r of the nominal model
    for s in range(8):
        Mr = nm.rnea(q[s], np.zeros(7), np.zeros(7), r[s]) - nm.rnea(q[s], np.zeros(7), np.zeros(7), np.zeros(7))
        assert Vs[s] >= 0.5 * r[s] @ Mr - 1e-12
    # the reference nominal torque is RNEA(q, q_d, qa_d, qa_dd)
    for s in range(8):
        qa_d = qd_d[s] + KR * (qd[s] - q[s])
        qa_dd = qd_dd[s] + KR * (qd_d[s] - q_d[s])
        assert np.abs(un[s] - nm.rnea(q[s], q_d[s], qa_d, qa_dd)).max() < 1e-11
    # below the threshold the robust input is zero
    u2, un2, v2, _, Vs2, _ = o.update(KR, ALPHA, V_MAX, 1e9, q, q_d, qd, qd_d, qd_dd)
    assert np.all(v2 == 0) and np.array_equal(u2, un2) and np.all(Vs2 == 0)
    # ALTHOFF method: v = -(Kp[1] ||bound|| + Kp[0]) r
    u3, un3, v3, ui3, _, _ = o.update_althoff(KR, KP, KI, MAX_ERR, q, q_d, qd, qd_d, qd_dd)
    assert np.array_equal(un3, un) and np.array_equal(ui3, ui)
    assert close(v3, -(KP[1] * np.linalg.norm(bound, axis=1) + KP[0])[:, None] * r)


def test_position_error_is_wrapped_to_pi():
    """clamp() of the position error (KRC/robust_controller.hpp:11-16): a full turn of the desired angle changes nothing
    but rounding."""
    o = _oracle.OracleController(MODEL)
    q, q_d, qd, qd_d, qd_dd = states(4, 4)
    a = o.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd)
    b = o.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd + 2 * np.pi, qd_d, qd_dd)
    assert close(a[0], b[0], 1e-9)


def test_oracle_matches_committed_golden():
    g = np.load(GOLDEN)
    o = _oracle.OracleController(MODEL)
    u, un, v, ui, Vs, outside = o.update(KR, ALPHA, V_MAX, R_THR, g["q"], g["q_d"], g["qd"], g["qd_d"], g["qd_dd"])
    assert outside == 0
    for name, val in (("u", u), ("u_nominal", un), ("v", v), ("u_interval", ui), ("V_sup", Vs)):
        assert close(val, g[name], 1e-12), name


def test_create_rejects_bad_model_files_without_a_gpu(tmp_path):
    import armour_b200 as ab
    L = ab.lib()
    h = ctypes.c_void_p()
    L.armour_controller_create.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    assert L.armour_controller_create(b"/nonexistent/robot.txt", 0.03, 0, ctypes.byref(h)) == -1
    assert b"could not open" in L.armour_last_error()
    text = open(MODEL).read().replace("parent <-1 0 1 2 3 4 5>", "parent <-1 0 1 1 3 4 5>")
    p = tmp_path / "tree.txt"
    p.write_text(text)
    assert L.armour_controller_create(str(p).encode(), 0.03, 0, ctypes.byref(h)) == -1
    assert b"serial chains" in L.armour_last_error()
    assert L.armour_controller_create(MODEL.encode(), 1.5, 0, ctypes.byref(h)) == -1


# ------------------------------------------------------------------------------------------------- device (GPU)
@pytest.fixture(scope="module")
def device_controller():
    from armour_b200.controller import RobustController
    c = RobustController(MODEL, 0.03, device=0)
    yield c
    c.close()


@pytest.mark.gpu
def test_device_update_matches_oracle(device_controller):
    o = _oracle.OracleController(MODEL, num_threads=8)
    q, q_d, qd, qd_d, qd_dd = states(10, 4096)
    u, un, v, ui, Vs, outside = device_controller.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd, debug=True)
    uo, uno, vo, uio, Vso, outside_o = o.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd)
    assert outside == 0 and outside_o == 0
    assert close(un, uno) and close(ui, uio) and close(Vs, Vso) and close(v, vo) and close(u, uo)
    # enclosure property on the device output itself
    assert np.all(ui[..., 0] <= un) and np.all(un <= ui[..., 1])


@pytest.mark.gpu
def test_device_althoff_matches_oracle(device_controller):
    o = _oracle.OracleController(MODEL, num_threads=8)
    q, q_d, qd, qd_d, qd_dd = states(11, 1000)
    u, un, v, ui, outside = device_controller.update_althoff(KR, KP, KI, MAX_ERR, q, q_d, qd, qd_d, qd_dd, debug=True)
    uo, uno, vo, uio, _, _ = o.update_althoff(KR, KP, KI, MAX_ERR, q, q_d, qd, qd_d, qd_dd)
    assert outside == 0
    assert close(un, uno) and close(ui, uio) and close(v, vo) and close(u, uo)


@pytest.mark.gpu
def test_device_single_tick_like_the_mex(device_controller):
    """count = 1 is the reference's call; a batch is the same computation per sample."""
    q, q_d, qd, qd_d, qd_dd = states(12, 33)
    U, UN, V = device_controller.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd)
    for s in (0, 17, 32):
        u, un, v = device_controller.update(KR, ALPHA, V_MAX, R_THR, q[s], q_d[s], qd[s], qd_d[s], qd_dd[s])
        assert np.array_equal(u, U[s]) and np.array_equal(un, UN[s]) and np.array_equal(v, V[s])
    # large tracking error, wrapped desired angle, threshold above ||r||
    o = _oracle.OracleController(MODEL)
    q2, q_d2, qd2, qd_d2, qd_dd2 = states(13, 64, err=1.0)
    qd2 = qd2 + 2 * np.pi
    a = device_controller.update(KR, ALPHA, V_MAX, R_THR, q2, q_d2, qd2, qd_d2, qd_dd2)
    b = o.update(KR, ALPHA, V_MAX, R_THR, q2, q_d2, qd2, qd_d2, qd_dd2)
    assert close(a[0], b[0]) and close(a[2], b[2])
    a = device_controller.update(KR, ALPHA, V_MAX, 1e9, q2, q_d2, qd2, qd_d2, qd_dd2)
    assert np.all(a[2] == 0) and np.array_equal(a[0], a[1])


@pytest.mark.gpu
def test_device_rnea_matches_oracle_and_newton_euler(device_controller):
    o = _oracle.OracleController(MODEL)
    rng = np.random.default_rng(14)
    q, qd, qda, qdd = (rng.uniform(-1, 1, (50, 7)) * s for s in (3.0, 1.0, 1.0, 2.0))
    tau = device_controller.rnea(q, qd, qda, qdd)
    ti = device_controller.rnea(q, qd, qda, qdd, interval=True)
    Mr = device_controller.rnea(q, 0 * qd, 0 * qd, qdd, gravity=False, interval=True)
    for s in range(50):
        assert np.abs(tau[s] - nm.rnea(q[s], qd[s], qda[s], qdd[s])).max() < 1e-11
        assert close(tau[s], o.rnea(q[s], qd[s], qda[s], qdd[s]))
        assert close(ti[s], o.rnea(q[s], qd[s], qda[s], qdd[s], interval=True))
        assert close(Mr[s], o.rnea(q[s], 0 * qd[s], 0 * qd[s], qdd[s], gravity=False, interval=True))


@pytest.mark.gpu
def test_device_resident_path_and_golden(device_controller):
    g = np.load(GOLDEN)
    device_controller.upload(g["q"], g["q_d"], g["qd"], g["qd_d"], g["qd_dd"])
    device_controller.update_resident(KR, ALPHA, V_MAX, R_THR)
    u, un, v, outside = device_controller.download()
    assert outside == 0 and device_controller.last_ms() > 0
    assert close(u, g["u"]) and close(un, g["u_nominal"]) and close(v, g["v"])


@pytest.mark.gpu
def test_device_large_batch_pipeline_equals_small_batches(device_controller):
    """count >= 32768 runs as a two-stream pipeline of 32768-tick chunks over page-locked caller arrays; the result must
    be the per-sample result of the plain path (70 000 ticks: two full chunks and a partial one)."""
    q, q_d, qd, qd_d, qd_dd = states(15, 70000)
    out = tuple(np.zeros((70000, 7)) for _ in range(3))
    for _ in range(2):   # second call reuses the registered buffers
        u, un, v = device_controller.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd, out=out)
    device_controller.release_host_buffers()
    for lo in (0, 32768 - 3, 65536 - 5, 70000 - 10):
        sl = slice(lo, lo + 10)
        us, uns, vs = device_controller.update(KR, ALPHA, V_MAX, R_THR, q[sl], q_d[sl], qd[sl], qd_d[sl], qd_dd[sl])
        assert np.array_equal(us, u[sl]) and np.array_equal(uns, un[sl]) and np.array_equal(vs, v[sl])
    # debug outputs through the pipeline as well
    r = device_controller.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd, debug=True)
    assert np.array_equal(r[0], u) and r[5] == 0
    assert np.all(r[3][..., 0] <= r[1]) and np.all(r[1] <= r[3][..., 1])
