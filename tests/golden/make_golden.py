"""Generates tests/golden/*.npz from the CPU oracle (oracle/).  The reference itself cannot run in this image
(Boost, Eigen, Ipopt, MATLAB are absent) and ships no golden vectors, so these fixtures pin the ORACLE's output
(regression guard) and give the GPU tests a committed target; they are not reference-generated.
Run:  python tests/golden/make_golden.py"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import _oracle
from problems import DEBUG_K, DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, EXAMPLE_OBS, EXAMPLE_Q0


def case(name, T, q0, qd0, qdd0, obs, xs):
    o = _oracle.Oracle(T=T)
    o.build(q0, qd0, qdd0, obs)
    data = dict(T=T, q0=q0, qd0=qd0, qdd0=qdd0, obs=np.asarray(obs, dtype=float), xs=np.array(xs), torque_radius=o.torque_radius(),
                link_generators=o.link_generators(), g=np.array([o.eval_g(x) for x in xs]), jac=np.array([o.eval_jac_g(x) for x in xs]))
    keys, counts = [], []
    for name_t in ("links", "u_nom"):
        for s in range(T):
            for i in range(7):
                z = o.get_pz(name_t, i, s)
                counts.append(len(z["keys"]))
                keys.extend(int(k) for k in z["keys"])
    data["k_only_counts"] = np.array(counts, dtype=np.int32)
    data["k_only_keys"] = np.array(keys, dtype=np.uint64)
    data["u_nom_centers"] = np.array([[o.get_pz("u_nom", i, s)["center"][0] for i in range(7)] for s in range(T)])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print(name, "m =", o.m, "k-only keys", len(keys))


if __name__ == "__main__":
    # state of KPR/debug_script.m:29-31 at the slice point of KPR/PZ_tests.cu:198, coarse grid to keep the file small
    case("debug_state_T16", 16, DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, EXAMPLE_OBS.reshape(-1, 12)[:3].ravel(), [DEBUG_K, np.zeros(7)])
    # example input of KPR/armour_main.cu:19-34 (zero initial velocity / acceleration)
    case("example_input_T16", 16, EXAMPLE_Q0, np.zeros(7), np.zeros(7), EXAMPLE_OBS.reshape(-1, 12)[:4].ravel(), [DEBUG_K, -DEBUG_K])
