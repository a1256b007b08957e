"""N > 1 host logic on CPU: world_size 2 over gloo (the GPU box runs the same code over NCCL)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from armour_b200 import sweep  # noqa: E402


def test_shard_is_a_partition():
    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            parts = [sweep.shard(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def oracle_solver(indices):
    """Per-problem record computed with the CPU oracle on a tiny configuration (test infrastructure)."""
    import _oracle
    from problems import make_problem, DEBUG_K
    out = np.zeros((len(indices), sweep.RECORD_WIDTH))
    for row, i in enumerate(indices):
        q0, qd0, qdd0, _, obs = make_problem(i, 2)
        o = _oracle.Oracle(T=4, num_threads=1)
        ms = o.build(q0, qd0, qdd0, obs)
        g = o.eval_g(DEBUG_K)
        out[row, :7] = g[:7]
        out[row, 7] = float(o.check_feasible(g))
        out[row, 8] = ms
        out[row, 11] = i
    return out


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = sweep.run_sweep(n, oracle_solver, rank=rank, world=world, device="cpu", batch=2)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_two_ranks_gloo_matches_single_process():
    n, world = 5, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = sweep.run_sweep(n, oracle_solver, rank=0, world=1)
    drop_timing = [c for c in range(sweep.RECORD_WIDTH) if c != 8]
    for r in range(world):
        assert results[r].shape == (n, sweep.RECORD_WIDTH)
        assert np.array_equal(results[r][:, 11], np.arange(n))
        assert np.array_equal(results[r][:, drop_timing], single[:, drop_timing])
