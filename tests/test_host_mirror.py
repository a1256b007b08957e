"""The C++ host mirror (armtd_NLP adapter, PZsparse facade, armour_main CLI) against the ctypes path, and the
"same solver, same k" check that stands in for the Ipopt comparison (Ipopt is not installed)."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import _oracle
import armour_b200 as ab
from problems import DEBUG_K, EXAMPLE_OBS, EXAMPLE_Q0, EXAMPLE_QDES, make_problem

pytestmark = pytest.mark.gpu
PKG = ab.PKG_DIR


def test_cpp_adapter_and_facade_match_ctypes_path(gpu_lib):
    exe = os.path.join(PKG, "test_host")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    tok = line.split()
    val = {tok[i]: tok[i + 1] for i in range(1, len(tok) - 1) if re.match(r"^[A-Za-z]", tok[i]) and not re.match(r"^[A-Za-z]", tok[i + 1])}
    q0 = [0.6543, -0.0876, -0.4837, -1.2278, -1.5735, -1.0720, 0]
    qd0 = [0.1, -0.1, 0.2, 0.1, -0.2, 0.1, 0.05]
    qdd0 = [0.1, 0.2, -0.1, 0.3, 0.1, -0.2, 0.1]
    obs = EXAMPLE_OBS.reshape(-1, 12)[[0, 2]].ravel()
    p = ab.Planner(T=16)
    p.build(q0, qd0, qdd0, obs)
    g, J = p.eval_g_jac(DEBUG_K)
    assert int(val["n"]) == 7 and int(val["m"]) == p.m
    assert abs(float(val["gsum"]) - g.sum()) <= 1e-9 * max(1, abs(g.sum()))
    assert abs(float(val["jsum"]) - J.sum()) <= 1e-9 * max(1, abs(J.sum()))
    assert abs(float(val["f"]) - p.eval_f(EXAMPLE_QDES, 0.5, DEBUG_K)) <= 1e-9
    R, L, U = p.get_pz("R", 0, 3), p.get_pz("links", 2, 3), p.get_pz("u_nom", 1, 3)
    RL = p.pz_binary("mul", R, L)
    C = p.pz_binary("cross", L, RL)
    assert int(val["RLn"]) == len(RL["keys"]) and int(val["Cn"]) == len(C["keys"])
    assert int(val["Sn"]) == len(p.pz_binary("add", U, U)["keys"]) and int(val["Dn"]) == 0
    assert int(val["iters"]) >= 1


def _write_armour_in(d, q0, qd0, qdd0, q_des, obs):
    """written like uarmtd_planner.m:169-196 (%.10f)"""
    obs = np.asarray(obs).reshape(-1, 12)
    with open(os.path.join(d, "armour.in"), "w") as f:
        for vec in (q0, qd0, qdd0, q_des):
            f.write(" ".join("%.10f" % v for v in vec) + "\n")
        f.write("%d\n" % len(obs))
        for row in obs:
            f.write(" ".join("%.10f" % v for v in row) + "\n")


@pytest.mark.parametrize("binary", ["armour_main", "armour_main_ipopt_stub"])
def test_armour_main_cli_file_protocol(binary, gpu_lib):
    """Text protocol of KPR/armour_main.cu:47-79 (input) and :324-397 (five output files): EVERY value of every file is
    compared with the in-memory path at the file's precision (10 significant digits; 6 for the constraint file), including
    the 28 bound rows appended to the constraint file (:381-394).  Both builds of the CLI are exercised: the stand-in solver
    branch and the ARMOUR_HAVE_IPOPT branch (SmartPtr ownership, options, return statuses) against the in-test Ipopt
    stand-in of tests/ipopt_stub — the same Gauss-Newton iteration sits behind both, so k must agree too."""
    exe = os.path.join(PKG, binary)
    assert os.path.exists(exe), "run __graft_entry__.build()"
    T, n_obs = 128, 10
    q0 = np.array([float("%.10f" % v) for v in EXAMPLE_Q0])      # what the executable reads back from the text file
    q_des = np.array([float("%.10f" % v) for v in EXAMPLE_QDES])
    obs = np.array([float("%.10f" % v) for v in EXAMPLE_OBS])
    with tempfile.TemporaryDirectory() as d:
        _write_armour_in(d, q0, np.zeros(7), np.zeros(7), q_des, obs)
        out = subprocess.run([exe, d], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        for line in ("Time taken by generating reachable sets", "Time allocated for Ipopt: 10000 milliseconds", "Time taken by Ipopt"):
            assert line in out.stdout, out.stdout
        assert ("Found an optimal solution!" in out.stdout) != ("Problem infeasible!" in out.stdout)
        m = 7 * T + 7 * T * n_obs + 28
        # the in-memory path: same build, same solver
        p = ab.Planner(T=T)
        p.build(q0, np.zeros(7), np.zeros(7), obs)
        k, feasible, _, _ = p.standin_solve(q_des, 0.5)
        g = p.eval_g(k)
        main = open(os.path.join(d, "armour.out")).read().split()
        if feasible:
            assert len(main) == 8 and "Found an optimal solution!" in out.stdout
            assert np.allclose([float(v) for v in main[:7]], k, rtol=1e-9, atol=1e-12)
        else:
            assert len(main) == 2 and main[0] == "-1" and "Problem infeasible!" in out.stdout
        assert float(main[-1]) >= 0 and float(main[-1]) == int(float(main[-1]))     # total milliseconds
        centers = np.loadtxt(os.path.join(d, "armour_joint_position_center.out"))
        assert centers.shape == (T * 7, 3)
        assert np.allclose(centers, p.link_sliced_center().reshape(T * 7, 3), rtol=1e-9, atol=1e-12)
        gens = np.loadtxt(os.path.join(d, "armour_joint_position_radius.out"))
        assert gens.shape == (T * 7 * 3, 6)
        assert np.allclose(gens, p.link_generators().reshape(T * 7 * 3, 6), rtol=1e-9, atol=1e-15)
        radius = np.loadtxt(os.path.join(d, "armour_control_input_radius.out"))
        assert radius.shape == (T, 7) and np.allclose(radius, p.torque_radius(), rtol=1e-9)
        cons = np.loadtxt(os.path.join(d, "armour_constraints.out"))
        assert cons.shape == (m + 28,)
        assert np.allclose(cons[:m], g, rtol=1e-5, atol=1e-12)
        _, _, gl, gu = p.get_bounds_info()
        pos = np.stack([gl[m - 28:m - 21], gu[m - 28:m - 21]], axis=1).ravel()      # lb, ub interleaved per joint (:382-386)
        vel = np.stack([gl[m - 14:m - 7], gu[m - 14:m - 7]], axis=1).ravel()        # (:389-393)
        assert np.allclose(cons[m:m + 14], pos, rtol=1e-5) and np.allclose(cons[m + 14:], vel, rtol=1e-5)
        # missing input file: -1 in armour.out and a non-zero exit code
        with tempfile.TemporaryDirectory() as d2:
            bad = subprocess.run([exe, d2], capture_output=True, text=True, timeout=60)
            assert bad.returncode != 0 and open(os.path.join(d2, "armour.out")).read().strip() == "-1"
        # too many obstacles (KPR/armour_main.cu:66-72)
        with tempfile.TemporaryDirectory() as d3:
            _write_armour_in(d3, q0, np.zeros(7), np.zeros(7), q_des, np.zeros(41 * 12))
            bad = subprocess.run([exe, d3], capture_output=True, text=True, timeout=60)
            assert bad.returncode != 0 and open(os.path.join(d3, "armour.out")).read().strip() == "-1"


def test_armour_main_ipopt_return_statuses(gpu_lib):
    """KPR/armour_main.cu:276-281, 295-317: initialisation failure, missing HSL library (Invalid_Option) and a CPU-time-out
    with a usable iterate, simulated by the in-test Ipopt stand-in (IPOPT_STUB_STATUS)."""
    exe = os.path.join(PKG, "armour_main_ipopt_stub")
    assert os.path.exists(exe)
    with tempfile.TemporaryDirectory() as d:
        _write_armour_in(d, EXAMPLE_Q0, np.zeros(7), np.zeros(7), EXAMPLE_QDES, EXAMPLE_OBS[:36])
        run = lambda status: subprocess.run([exe, d], capture_output=True, text=True, timeout=300, env=dict(os.environ, IPOPT_STUB_STATUS=status, ARMOUR_NUM_TIME_STEPS="16"))
        r = run("Maximum_CpuTime_Exceeded")
        assert r.returncode == 0 and "Ipopt maximum CPU time exceeded!" in r.stdout and "Time taken by Ipopt" in r.stdout
        assert ("Found a feasible solution!" in r.stdout) != ("Did not find a feasible solution!" in r.stdout)
        main = open(os.path.join(d, "armour.out")).read().split()
        assert len(main) == (8 if "Found a feasible solution!" in r.stdout else 2)
        r = run("Invalid_Option")
        assert r.returncode == 0 and "Cannot find HSL library!" in r.stdout and "Time taken by Ipopt" not in r.stdout
        assert open(os.path.join(d, "armour.out")).read().split()[0] == "-1"
        r = run("Initialize_Failure")
        assert r.returncode != 0 and "Error during initialization!" in r.stdout
        assert open(os.path.join(d, "armour.out")).read().strip() == "-1"


def test_same_solver_same_k(gpu_lib):
    """north_star: 'the Ipopt-returned k within 1e-6'.  Ipopt is not installed, so the claim is checked as: the same
    host solver (scipy SLSQP, a stand-in) driving the oracle's and the device's callbacks returns the same k."""
    from scipy.optimize import minimize
    T, n_obs = 16, 4
    q0, qd0, qdd0, q_des, obs = make_problem(31, n_obs)
    o = _oracle.Oracle(T=T)
    o.build(q0, qd0, qdd0, obs)
    p = ab.Planner(T=T)
    p.build(q0, qd0, qdd0, obs)

    def solve(b):
        _, _, gl, gu = b.get_bounds_info()
        lo, hi = gl > -1e18, gu < 1e18

        def cons(x):
            g = b.eval_g(x)
            return np.concatenate([(g - gl)[lo], (gu - g)[hi]])

        def jac(x):
            J = b.eval_jac_g(x)
            return np.concatenate([J[lo], -J[hi]])

        r = minimize(lambda x: b.eval_f(q_des, 0.5, x), np.zeros(7), jac=lambda x: b.eval_grad_f(q_des, 0.5, x), bounds=[(-1, 1)] * 7,
                     constraints=[{"type": "ineq", "fun": cons, "jac": jac}], method="SLSQP", options={"maxiter": 200, "ftol": 1e-12})
        return r.x

    k_o, k_g = solve(o), solve(p)
    assert np.abs(k_o - k_g).max() <= 1e-6, (k_o, k_g)
