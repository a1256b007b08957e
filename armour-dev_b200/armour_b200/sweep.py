"""Batched sweep of independent planning problems across ranks.

Problems are fully independent, so the path shards by problem index with no data-path collective;
the only exchange is ONE all_gather of fixed-size per-problem result records at the end.  Works with
any torch.distributed backend: NCCL on the GPU box (one rank per B200), gloo in the CPU tests.
"""
import numpy as np

RECORD_WIDTH = 12   # k[7], feasible, build_ms, solve_ms, n_iter, problem index


def shard(n_problems, rank, world):
    """Block partition [lo, hi) of problem indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_problems, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pad_to(n_problems, world):
    return (n_problems + world - 1) // world


def run_sweep(n_problems, solve_fn, rank=0, world=1, device="cpu", batch=64):
    """solve_fn(indices: list[int]) -> float array [len(indices), RECORD_WIDTH].  Returns the gathered
    [n_problems, RECORD_WIDTH] array (identical on every rank), rows ordered by problem index."""
    import torch
    import torch.distributed as dist

    lo, hi = shard(n_problems, rank, world)
    per = pad_to(n_problems, world)
    local = np.full((per, RECORD_WIDTH), np.nan)
    for b0 in range(lo, hi, batch):
        idx = list(range(b0, min(b0 + batch, hi)))
        rec = np.asarray(solve_fn(idx), dtype=np.float64)
        assert rec.shape == (len(idx), RECORD_WIDTH)
        local[b0 - lo: b0 - lo + len(idx)] = rec
    if world == 1:
        return local[: hi - lo]
    mine = torch.from_numpy(local).to(device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)          # the only collective on the path
    out = np.empty((n_problems, RECORD_WIDTH))
    for r in range(world):
        rlo, rhi = shard(n_problems, r, world)
        out[rlo:rhi] = gathered[r].cpu().numpy()[: rhi - rlo]
    return out
