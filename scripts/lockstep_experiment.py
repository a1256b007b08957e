"""Is the sweep limited by instruction fetch?  With T = 148 intervals and a static work stride (4 x 148 CTAs), the four CTAs
resident on one SM process the SAME interval index of consecutive problems.  If all problems of the batch are identical the
four CTAs execute identical instruction streams in near lockstep (shared instruction-cache lines); with different problems
they run the same program on different data (different sizes, drifting apart).  Same total work per problem otherwise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import armour_b200 as ab
from problems import make_problem
T, B, n_obs = 148, 64, 4
for label, seeds in (("identical problems (lockstep)", [5000] * B), ("different problems", [5000 + i for i in range(B)]), ("identical problems (lockstep)", [5007] * B)):
    pb = ab.Planner(T=T, max_obstacles=n_obs, device=0, batch=B)
    bp = [make_problem(s, n_obs) for s in seeds]
    pb.upload_problems(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]), np.concatenate([q[4] for q in bp]), n_obs)
    pb.build_resident()
    ms = []
    for _ in range(3):
        pb.build_resident(); ms.append(pb.last_build_ms()[1])
    print("%-32s %.2f ms per batch of %d x %d intervals  (%.1f us per interval-CTA slot)" % (label, np.mean(ms), B, T, np.mean(ms) * 1e3 * 592 / (B * T)), flush=True)
    pb.close()
