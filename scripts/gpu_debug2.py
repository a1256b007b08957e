import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np
from _oracle import Oracle
import armour_b200 as ab
from problems import make_problem
q0, qd0, qdd0, q_des, obs = make_problem(0, 10)
o = Oracle(T=128); o.build(q0, qd0, qdd0, obs)
p = ab.Planner(T=128); p.build(q0, qd0, qdd0, obs)
(co, so), (cg, sg) = o.taylor_remainders(), p.taylor_remainders()
for name, ref, dev in (("cos", co, cg), ("sin", so, sg)):
    badlo = np.argwhere(dev[..., 0] > ref[..., 0]); badhi = np.argwhere(dev[..., 1] < ref[..., 1])
    print(name, "lo violations", len(badlo), "hi violations", len(badhi))
    for j, s in list(badlo[:6]):
        print("  lo", j, s, repr(dev[j, s, 0]), repr(ref[j, s, 0]), dev[j, s, 0] - ref[j, s, 0], "hi diff", dev[j, s, 1] - ref[j, s, 1])
    for j, s in list(badhi[:6]):
        print("  hi", j, s, repr(dev[j, s, 1]), repr(ref[j, s, 1]), dev[j, s, 1] - ref[j, s, 1], "lo diff", dev[j, s, 0] - ref[j, s, 0])
    print("  typical widening lo", np.median(ref[..., 0] - dev[..., 0]), "hi", np.median(dev[..., 1] - ref[..., 1]))
