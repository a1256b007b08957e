"""Writes tests/golden/controller_64.npz: 64 seeded controller ticks and the oracle's outputs for them
(oracle/oracle_controller.cpp, ARMOUR robust input, defaults of uarmtd_robust_CBF_LLC.m:6-9).  Regression fixture: it
pins the oracle against accidental change and gives the device path a fixed target on the GPU box."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import _oracle
from test_controller import states, MODEL, KR, ALPHA, V_MAX, R_THR

q, q_d, qd, qd_d, qd_dd = states(2024, 64)
o = _oracle.OracleController(MODEL)
u, un, v, ui, Vs, outside = o.update(KR, ALPHA, V_MAX, R_THR, q, q_d, qd, qd_d, qd_dd)
assert outside == 0
np.savez(os.path.join(HERE, "controller_64.npz"), q=q, q_d=q_d, qd=qd, qd_d=qd_d, qd_dd=qd_dd, u=u, u_nominal=un, v=v, u_interval=ui, V_sup=Vs)
print("wrote controller_64.npz")
