"""Diagnostic: armour_eval_batch with page-locked caller arrays from several processes at once (world size from torchrun or 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
from problems import make_problem
rank = int(os.environ.get("LOCAL_RANK", "0")); dev = rank if os.environ.get("DIAG_MULTI_GPU") else 0
B, T, n_obs = 256, 128, 10
pb = ab.Planner(T=T, max_obstacles=n_obs, device=dev, batch=B, pin_user_buffers=True)
m = 7 * T + 7 * T * n_obs + 28
G, V = np.zeros((B, m)), np.zeros((B, m * 7))
probs = [make_problem(i, n_obs) for i in range(8)] * (B // 8)
pb.build_batch(*[np.concatenate([q[k] for q in probs]) for k in (0, 1, 2, 4)], n_obs)
try:
    pb.eval_batch(np.zeros((B, 7)), g=G)
    pb.eval_batch(np.zeros((B, 7)), g=G, values=V)
    print("rank", rank, "ok", float(G.sum()), float(V.sum()), flush=True)
except Exception as e:
    print("rank", rank, "FAILED", e, flush=True)
