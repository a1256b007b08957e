"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): one T = 8 single-plan build (two thread groups
per CTA with shared-memory hand-off counters), one 2-problem batch in the narrow sweep shape, one fused constraint evaluation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
from problems import make_problem, DEBUG_K
T = int(os.environ.get("SAN_T", "8"))
q0, qd0, qdd0, _, obs = make_problem(123, 3)
p = ab.Planner(T=T, device=0)
p.build(q0, qd0, qdd0, obs)
g, J = p.eval_g_jac(DEBUG_K)
print("single ok", float(np.abs(g).sum()))
p.close()
pb = ab.Planner(T=T, device=0, batch=2)
bp = [make_problem(5 + i, 3) for i in range(2)]
pb.build_batch(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]), np.concatenate([q[4] for q in bp]), 3)
g, J = pb.eval_g_jac(DEBUG_K)
print("batch ok", float(np.abs(g).sum()))
pb.close()
