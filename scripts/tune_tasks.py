"""Task-scheduled reach kernel: timing and a quick parity check per configuration (ARMOUR_TUNE_TASKS = thread groups)."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]


def worker():
    import armour_b200 as ab
    import _oracle
    from problems import make_problem
    p = ab.Planner(T=128, max_obstacles=20, device=0)
    ms = []
    for s in range(4):
        q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
        p.build(q0, qd0, qdd0, obs)
        ms.append(p.last_build_ms()[1])
    o = _oracle.Oracle(T=128, num_threads=16)
    o.build(q0, qd0, qdd0, obs)
    bad = 0
    for name in ("links", "u_nom"):
        for t in range(0, 128, 9):
            for j in range(7):
                a, b = o.get_pz(name, j, t), p.get_pz(name, j, t)
                if not np.array_equal(a["keys"], b["keys"]) or np.abs(a["coeffs"] - b["coeffs"]).max(initial=0) > 1e-12:
                    bad += 1
    tr = np.abs(p.torque_radius() - o.torque_radius()).max()
    print("RESULT reach_ms %s  mismatching tables %d  max|torque radius diff| %.2e" % (" ".join("%.3f" % m for m in ms), bad, tr), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        worker()
    else:
        for tasks, scap, tcap in ((0, 0, 0), (4, 1536, 256), (4, 1792, 256), (4, 1280, 256), (4, 1792, 128)):
            env = dict(os.environ, ARMOUR_TUNE_TASKS=str(tasks), ARMOUR_TUNE_TASK_SCAP=str(scap or 1024), ARMOUR_TUNE_TASK_TCAP=str(tcap or 512))
            try:
                out = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True, timeout=120)
                res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
                print("tasks=%d scap=%d tcap=%d:" % (tasks, scap, tcap), res[0] if res else "FAILED " + out.stderr[-400:], flush=True)
            except subprocess.TimeoutExpired:
                print("tasks=%d scap=%d tcap=%d: TIMEOUT" % (tasks, scap, tcap), flush=True)
