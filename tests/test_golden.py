"""Committed golden fixtures (tests/golden/*.npz, generated from the oracle by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the device path must match them (GPU)."""
import glob
import os

import numpy as np
import pytest

import _oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith(("controller_", "worlds_")))   # controller fixtures: tests/test_controller.py


def k_only(backend, T):
    counts, keys = [], []
    for name in ("links", "u_nom"):
        for s in range(T):
            for i in range(7):
                z = backend.get_pz(name, i, s)
                counts.append(len(z["keys"]))
                keys.extend(int(k) for k in z["keys"])
    return np.array(counts, dtype=np.int32), np.array(keys, dtype=np.uint64)


def check(backend, d, tol_reach, tol_con, radii_one_sided):
    T = int(d["T"])
    counts, keys = k_only(backend, T)
    assert np.array_equal(counts, d["k_only_counts"]) and np.array_equal(keys, d["k_only_keys"])      # bit-exact keys
    cen = np.array([[backend.get_pz("u_nom", i, s)["center"][0] for i in range(7)] for s in range(T)])
    assert np.abs(cen - d["u_nom_centers"]).max() <= tol_reach * np.abs(d["u_nom_centers"]).max()
    tr = backend.torque_radius()
    assert np.abs(tr - d["torque_radius"]).max() <= tol_reach * np.abs(tr).max()
    if radii_one_sided:
        assert np.all(tr >= d["torque_radius"])
    assert np.abs(backend.link_generators() - d["link_generators"]).max() <= tol_reach
    for x, g, J in zip(d["xs"], d["g"], d["jac"]):
        assert np.abs(backend.eval_g(x) - g).max() <= tol_con * max(1.0, np.abs(g).max())
        assert np.abs(backend.eval_jac_g(x) - J).max() <= tol_con * max(1.0, np.abs(J).max())


def test_fixtures_exist():
    assert len(FIXTURES) >= 2


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_golden(path):
    d = np.load(path)
    o = _oracle.Oracle(T=int(d["T"]))
    o.build(d["q0"], d["qd0"], d["qdd0"], d["obs"])
    check(o, d, 1e-13, 1e-12, False)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_device_matches_golden(path, gpu_lib):
    import armour_b200 as ab
    d = np.load(path)
    p = ab.Planner(T=int(d["T"]))
    p.build(d["q0"], d["qd0"], d["qdd0"], d["obs"])
    check(p, d, 1e-9, 1e-8, True)
