"""Config 3 (BASELINE.json): batched sweep of independent planning problems sharded over 1/2/4/8 GPUs with ONE
all_gather of per-problem result records.  python scripts/run_sweep.py --problems 4096  (torchrun for N > 1).
Per problem: reach-set + constraint build on the device (batched launches) and a host-driven solve with the
stand-in solver (Ipopt is not installed; the solve is host-bound and reported separately from the builds)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--n-obs", type=int, default=10)
    ap.add_argument("--no-solve", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    import armour_b200 as ab
    from armour_b200 import sweep
    from problems import make_problem
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pb = ab.Planner(T=128, max_obstacles=args.n_obs, device=local_rank, batch=args.batch)
    stats = {"build_ms": 0.0, "solve_s": 0.0, "evals": 0}

    def solve_fn(indices):
        probs = [make_problem(i, args.n_obs) for i in indices]
        ms = pb.build_batch(np.concatenate([p[0] for p in probs]), np.concatenate([p[1] for p in probs]), np.concatenate([p[2] for p in probs]),
                            np.concatenate([p[4] for p in probs]), args.n_obs)
        stats["build_ms"] += ms
        out = np.zeros((len(indices), sweep.RECORD_WIDTH))
        for row, (i, p) in enumerate(zip(indices, probs)):
            out[row, 8] = ms / len(indices)
            out[row, 11] = i
            if args.no_solve:
                continue
            pb.select_problem(row)
            t0 = time.perf_counter()
            k, feas, it, ev = pb.standin_solve(p[3], 0.5)
            dt = time.perf_counter() - t0
            stats["solve_s"] += dt; stats["evals"] += ev
            out[row, :7] = k; out[row, 7] = float(feas); out[row, 9] = dt * 1e3; out[row, 10] = it
        return out

    solve_fn(list(range(min(args.batch, 4))))   # warm-up
    stats.update(build_ms=0.0, solve_s=0.0, evals=0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sweep.run_sweep(args.problems, solve_fn, rank=rank, world=world, device="cuda", batch=args.batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    agg = torch.tensor([stats["build_ms"], stats["solve_s"], float(stats["evals"])], dtype=torch.float64, device="cuda")
    mx = agg.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "config": "batched sweep, T=128, %d obstacles, %d problems over %d GPU(s), batch %d" % (args.n_obs, args.problems, world, args.batch),
            "wall_s": wall, "problems_per_s_incl_host_solve": args.problems / wall,
            "device_build_ms_max_over_ranks": float(mx[0]), "builds_per_s_device_only": args.problems / (float(mx[0]) / 1e3),
            "host_solve_s_max_over_ranks": float(mx[1]), "constraint_evals_rank0": int(stats["evals"]),
            "feasible_fraction": float(np.nanmean(res[:, 7])) if not args.no_solve else None,
            "solver": "stand-in (Ipopt not installed)" if not args.no_solve else "none", "records_checksum": float(np.nansum(res[:, :8])),
        }), flush=True)
    pb.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
