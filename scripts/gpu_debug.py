"""First-contact GPU check: build one problem on the device and on the oracle and print the differences."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
from _oracle import Oracle
import armour_b200 as ab
from problems import make_problem, DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0, DEBUG_K

T = int(os.environ.get("T", "128"))
n_obs = int(os.environ.get("NOBS", "10"))
q0, qd0, qdd0, q_des, obs = make_problem(3, n_obs)
if os.environ.get("DEBUGSTATE"):
    q0, qd0, qdd0 = DEBUG_Q0, DEBUG_QD0, DEBUG_QDD0
o = Oracle(T=T); t = time.time(); o.build(q0, qd0, qdd0, obs); print("oracle build s", time.time() - t, "threads", o.L.oracle_num_threads())
p = ab.Planner(T=T, threads_per_cta=int(os.environ.get("NT", "256")))
for rep in range(3):
    ms = p.build(q0, qd0, qdd0, obs); print("gpu build ms", p.last_build_ms())

def cmp_pz(name, idx, s):
    a, b = o.get_pz(name, idx, s), p.get_pz(name, idx, s)
    if len(a["keys"]) != len(b["keys"]) or not np.array_equal(a["keys"], b["keys"]):
        return ("KEYS", len(a["keys"]), len(b["keys"]))
    sc = max(1.0, np.abs(a["center"]).max())
    e = [np.abs(a["center"] - b["center"]).max() / sc]
    if len(a["keys"]): e.append(np.abs(a["coeffs"] - b["coeffs"]).max())
    ri = np.abs(a["independent"] - b["independent"]).max() / max(1e-300, np.abs(a["independent"]).max())
    return ("ok", max(e), ri, bool(np.all(b["independent"] >= a["independent"])))

for name in ("cos_q", "sin_q", "R", "R_t", "qd_des", "qda_des", "qdda_des", "links", "u_nom", "u_nom_int"):
    bad = 0; worst = 0; worstr = 0; contained = True
    for s in range(T):
        for i in range(7):
            r = cmp_pz(name, i, s)
            if r[0] != "ok":
                bad += 1
                if bad < 4: print("  ", name, i, s, r)
            else:
                worst = max(worst, r[1]); worstr = max(worstr, r[2]); contained &= r[3]
    print("%-10s key mismatches %d  max coeff/centre err %.3e  max rel radius err %.3e  radii>=oracle %s" % (name, bad, worst, worstr, contained))
tr_o, tr_g = o.torque_radius(), p.torque_radius()
print("torque radius rel err", np.abs(tr_o - tr_g).max() / np.abs(tr_o).max(), "gpu>=oracle", bool(np.all(tr_g >= tr_o)))
lg_o, lg_g = o.link_generators(), p.link_generators()
print("link gens err", np.abs(lg_o - lg_g).max())
co, so = o.taylor_remainders(); cg, sg = p.taylor_remainders()
print("taylor containment", bool(np.all(cg[..., 0] <= co[..., 0]) and np.all(cg[..., 1] >= co[..., 1]) and np.all(sg[..., 0] <= so[..., 0]) and np.all(sg[..., 1] >= so[..., 1])),
      "max widening", max(np.abs(cg - co).max(), np.abs(sg - so).max()))
if n_obs:
    Ao, do_, dlo = o.hyperplanes(); Ag, dg, dlg = p.hyperplanes()
    print("hyperplanes err", np.abs(Ao - Ag).max(), np.abs(do_ - dg).max(), np.abs(dlo - dlg).max())
for x in (DEBUG_K, np.zeros(7), np.random.default_rng(1).uniform(-1, 1, 7)):
    go, Jo = o.eval_g(x), o.eval_jac_g(x)
    t = time.time(); gg, Jg = p.eval_g_jac(x); dt = time.time() - t
    print("eval: g err %.3e  jac err %.3e  (kernel %.3f ms, wall %.3f ms)" % (np.abs(go - gg).max(), np.abs(Jo - Jg).max(), p.last_eval_ms(), dt * 1e3))
    bad = np.argsort(-np.abs(Jo - Jg).max(axis=1))[:3]; print("   worst jac rows", bad, np.abs(Jo - Jg).max(axis=1)[bad])
print("bounds err", [np.abs(a - b).max() for a, b in zip(o.get_bounds_info(), p.get_bounds_info())])
print("launches", p.kernel_launches())
