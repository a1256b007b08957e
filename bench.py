#!/usr/bin/env python
"""bench.py — reach-set + constraint builds per second on BASELINE.json's config[1]
(Kinova Gen3 single plan, T = 128 intervals, 20 obstacles, one fused eval_g/eval_jac_g per build).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    # the reference itself (oracle/_ref, else the oracle port) on the host cores

A "step" is one pass of the hot path for one planning problem: stages A-D of the reference's main()
(KPR/armour_main.cu:89-226: joint reach sets, PZ forward kinematics, PZ-RNEA nominal+interval, torque radius,
half-space tables) followed by one evaluation of all constraints and their Jacobian at a point k
(KPR/NLPclass.cu:272-396).  Every step uses a new synthetic problem (tests/problems.py, seed-indexed).
`value` times the step with inputs already resident in HBM (CUDA events on the handle's own stream);
`e2e` times the same step through the public C ABI with host buffers, host<->device copies included.
N > 1: independent problems per rank (weak scaling), one NCCL all_gather of per-problem result records.
"""
import os as _os
import sys as _sys
# stdout carries the one JSON line only: everything libraries print while the bench runs (e.g. NCCL's version banner) goes
# to stderr; _emit() writes the line to the real stdout.
_REAL_STDOUT = _os.dup(1)
_os.dup2(2, 1)


def _emit(text):
    _sys.stdout.flush()
    _os.write(_REAL_STDOUT, (text + "\n").encode())


import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "armour-dev_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

T = 128
N_OBS = 20
WORKLOAD = "kinova_gen3_single_plan_T128_obs20: reach-set+constraint build (stages A-D) + 1 fused eval_g/eval_jac_g"
METRIC = "reach_set_builds_per_sec"
UNIT = "builds/s"


def problem_for(rank, step):
    from problems import make_problem
    return make_problem(100000 + 1000 * rank + step, N_OBS)


def x_for(rank, step):
    return np.random.default_rng(777 + 1000 * rank + step).uniform(-1, 1, 7)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the measured phase (profiling recipe's clocks line)."""

    def __init__(self, dev):
        self.dev, self.rows, self.proc, self.thread = dev, [], None, None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_full_summary.csv")
    out = {}
    try:
        import csv
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rows[2:]:
            name = "reach_build_kernel" if "reach_build" in r[0] else "constraint_eval_kernel" if "constraint_eval" in r[0] else None
            if name:
                out[name] = (float(r[ir]) + float(r[iw])) * 1e6   # reported in Mbyte
    except Exception:
        pass
    return out


def eval_algorithmic_bytes(m):
    """half-space table read once + sliceable tables + outputs, per eval_g+eval_jac_g pair."""
    table = 40 * 36 * T * 7 * N_OBS
    sliceable = 0.5e6
    outputs = 8 * m + 56 * m
    return table + sliceable + outputs


def reference_runner(cores):
    """Callable timing one step (build + eval_g + eval_jac_g) of the reference arm, plus how to describe it.
    Preferred: oracle/_ref/libref_cuda.so = the reference's OWN sources (PZsparse/Trajectory/Dynamics on the host cores with
    OpenMP, its CollisionChecking kernels and armtd_NLP callbacks) compiled unmodified against stand-in Eigen/Boost/Ipopt
    headers (oracle/Makefile target ref).  Fallback: the oracle port."""
    import torch
    import _oracle
    if os.path.exists(_oracle.REF_CUDA_LIB_PATH) and T == 128 and torch.cuda.is_available():
        ref = _oracle.ReferenceCuda(num_threads=cores)

        def step(q0, qd0, qdd0, obs, x):
            t0 = time.perf_counter()
            ref.build(q0, qd0, qdd0, q0, obs)
            ref.eval_g(x)
            ref.eval_jac_g(x)
            return time.perf_counter() - t0
        return step, "reference", ("the reference's own sources (oracle/_ref, stand-in Eigen/Boost/Ipopt headers): reach sets on %d host threads "
                                   "(OpenMP over time intervals, KPR/armour_main.cu:100,118), its half-space / plane-test kernels on the GPU as in the reference" % cores)
    o = _oracle.Oracle(T=T, num_threads=cores)

    def step(q0, qd0, qdd0, obs, x):
        t0 = time.perf_counter()
        o.build(q0, qd0, qdd0, obs)
        o.eval_g(x)
        o.eval_jac_g(x)
        return time.perf_counter() - t0
    return step, "port", "oracle port of the reference algorithm on %d host threads (OpenMP over time intervals like KPR/armour_main.cu:100,118)" % cores


def run_reference(args):
    """The reference on the box's host cores, all threads, same workload / metric / unit (see reference_runner)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))   # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    step_fn, kind, how = reference_runner(cores)
    times = []
    for step in range(args.warmup + args.steps):
        q0, qd0, qdd0, _, obs = problem_for(0, step)
        dt = step_fn(q0, qd0, qdd0, obs, x_for(0, step))
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = 1e3 / ms
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "time_intervals": T, "obstacles": N_OBS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d full steps (1 build + eval_g + eval_jac_g each); %s" % (args.steps, how)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(json.dumps(line))


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    import armour_b200 as ab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    p = ab.Planner(T=T, max_obstacles=N_OBS, device=local_rank, threads_per_cta=args.threads, pin_user_buffers=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    probs = [problem_for(rank, s) for s in range(W + K)]
    xs = [x_for(rank, s) for s in range(W + K)]
    m = 7 * T + 7 * T * N_OBS + 28
    g_host, jac_host = np.zeros(m), np.zeros(m * 7)

    for s in range(W):   # warm-up: module load, first-touch of the arena, capacity growth if any
        q0, qd0, qdd0, _, obs = probs[s]
        p.build(q0, qd0, qdd0, obs)
        p.eval_g_jac(xs[s], g_host, jac_host)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- value: inputs resident in HBM, device time from CUDA events on the handle's stream ----
    launches0 = p.kernel_launches()
    build_ms, reach_ms, hyper_ms, eval_ms = [], [], [], []
    records = torch.zeros((K, 4), dtype=torch.float64, device="cuda")
    barrier()
    t_wall0 = time.perf_counter()
    for s in range(W, W + K):
        q0, qd0, qdd0, _, obs = probs[s]
        p.upload_problems(q0, qd0, qdd0, obs, N_OBS)
        p.upload_x(xs[s])
        flush_l2()
        p.build_resident()
        p.eval_resident(None)
        b, r, h = p.last_build_ms()
        build_ms.append(b); reach_ms.append(r); hyper_ms.append(h); eval_ms.append(p.last_eval_ms())
    launches = p.kernel_launches() - launches0
    # fixed-size per-problem result records: seed, build ms, eval ms, feasible-at-k flag (host solve not included:
    # Ipopt is absent from this image)
    rec_host = np.array([[100000 + 1000 * rank + s, build_ms[s - W], eval_ms[s - W], 0.0] for s in range(W, W + K)])
    records.copy_(torch.from_numpy(rec_host))
    if world > 1:
        gathered = [torch.zeros_like(records) for _ in range(world)]
        dist.all_gather(gathered, records)   # the only data-path collective: per-problem result records
    barrier()
    wall_value_region = time.perf_counter() - t_wall0
    dev_ms = torch.tensor([float(np.sum(build_ms) + np.sum(eval_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(dev_ms.item()) / K
    value = world * 1e3 / ms_per_step

    # ---- e2e: the public C-ABI calls with host buffers, copies inside the timed region ----
    e2e_t = []
    barrier()
    for s in range(W, W + K):
        q0, qd0, qdd0, _, obs = probs[s]
        flush_l2()
        t0 = time.perf_counter()
        p.build(q0, qd0, qdd0, obs)
        p.eval_g_jac(xs[s], g_host, jac_host)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = torch.tensor([float(np.sum(e2e_t)) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(e2e_ms.item()) / K
    h2d = 21 * 8 + N_OBS * 12 * 8 + 7 * 8
    d2h = 8 * m + 56 * m + T * 7 * 8 + 4

    extra = {}
    if rank == 0:
        # per-iteration constraint latency (config 2): 100 random k after one build
        lat_wall, lat_kern = [], []
        rng = np.random.default_rng(1234)
        for _ in range(100):
            x = rng.uniform(-1, 1, 7)
            t0 = time.perf_counter()
            p.eval_g_jac(x, g_host, jac_host)
            lat_wall.append((time.perf_counter() - t0) * 1e6)
            lat_kern.append(p.last_eval_ms() * 1e3)
        extra["eval_g_jac_latency_us"] = {"p50": float(np.percentile(lat_wall, 50)), "p99": float(np.percentile(lat_wall, 99)),
                                          "kernel_p50": float(np.percentile(lat_kern, 50)), "calls": 100}
        extra["plan_step_latency_ms_p50"] = {"build_plus_1_eval_e2e": float(np.percentile(np.array(e2e_t) * 1e3, 50)),
                                             "note": "Ipopt is not installed in this image; a full solve adds (iterations x eval latency)"}
    # batched sweep throughput (config 3 shape): B independent problems in one launch
    if args.sweep_batch > 0:
        B = args.sweep_batch
        pb = ab.Planner(T=T, max_obstacles=N_OBS, device=local_rank, batch=B, threads_per_cta=args.threads)
        bp = [problem_for(rank, 5000 + i) for i in range(B)]
        pb.upload_problems(np.concatenate([q[0] for q in bp]), np.concatenate([q[1] for q in bp]), np.concatenate([q[2] for q in bp]),
                           np.concatenate([q[4] for q in bp]), N_OBS)
        pb.build_resident()
        sw = []
        for _ in range(3):
            flush_l2()
            pb.build_resident()
            sw.append(pb.last_build_ms()[0])
        sw_ms = torch.tensor([float(np.mean(sw))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sw_ms, op=dist.ReduceOp.MAX)
        extra_sweep = {"batch_per_gpu": B, "ms_per_batch": float(sw_ms.item()), "builds_per_s": world * B * 1e3 / float(sw_ms.item())}
        pb.close()
    else:
        extra_sweep = None
    clocks = sampler.stop()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fp64_peak = ab.measure_fp64_peak(local_rank)
        # algorithmic flops per build: the oracle's op counter on the reference op sequence
        import _oracle
        flops, cpu_t, port_t = [], [], []
        cores = len(os.sched_getaffinity(0))
        o = _oracle.Oracle(T=T, num_threads=cores)
        ref_step, ref_kind, ref_how = reference_runner(cores) if world == 1 else (None, None, None)
        n_sample = 0
        t_budget = time.perf_counter()
        for s in range(W, W + K):
            q0, qd0, qdd0, _, obs = probs[s]
            t0 = time.perf_counter()
            o.build(q0, qd0, qdd0, obs)
            o.eval_g(xs[s]); o.eval_jac_g(xs[s])
            port_t.append(time.perf_counter() - t0)
            flops.append(o.op_stats()["flops"])
            if ref_step is not None:
                cpu_t.append(ref_step(q0, qd0, qdd0, obs, xs[s]))
            n_sample += 1
            if time.perf_counter() - t_budget > 20.0:
                break
        flops_per_build = float(np.mean(flops))
        reach_mean_ms = float(np.mean(reach_ms))
        achieved = flops_per_build / (reach_mean_ms * 1e-3) / 1e12
        eval_bytes = eval_algorithmic_bytes(m)
        traffic = ncu_traffic()
        eval_mean_ms = float(np.mean(eval_ms))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "time_intervals": T, "obstacles": N_OBS, "l2": "256 MiB flush between timed steps",
                       "threads_per_cta": args.threads or 256, "parallelism": "1 plan per GPU, weak scaling over independent problems"},
            "e2e": {"value": world * 1e3 / e2e_ms_per_step, "unit": UNIT, "ms_per_step": e2e_ms_per_step, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "reach_build_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": traffic.get("reach_build_kernel"), "traffic_source": "profiles/r1_ncu_full_summary.csv (ncu --set full, one launch)", "algorithmic_flops_per_launch": flops_per_build, "kernel_ms": reach_mean_ms,
                         "peak_source": "fp64 FMA micro-benchmark in this run (MEASURED_PEAKS.json has no fp64 entry)",
                         "note": "a single plan is 128 CTAs of dependent small sorts: latency-bound, far from the fp64 roofline (SURVEY.md §7)"},
            "roofline_eval": {"bound": "hbm", "kernel": "constraint_eval_kernel", "achieved": eval_bytes / (eval_mean_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                              "unit": "GB/s", "frac": eval_bytes / (eval_mean_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": traffic.get("constraint_eval_kernel"),
                              "algorithmic_bytes_per_launch": eval_bytes, "kernel_ms": eval_mean_ms, "peak_source": peak_src},
            "kernel_ms": {"reach_build": reach_mean_ms, "hyperplanes": float(np.mean(hyper_ms)), "constraint_eval": eval_mean_ms},
            "wall_s_value_region": wall_value_region,
        }
        if world == 1:
            cpu_ms = 1e3 * float(np.mean(cpu_t))
            line["cpu_baseline"] = {"value": 1e3 / cpu_ms, "unit": UNIT, "cores": cores, "kind": ref_kind, "ms_per_step": cpu_ms,
                                    "sample": "%d of the timed steps (same problems), 1 build + eval_g + eval_jac_g each; %s" % (n_sample, ref_how),
                                    "oracle_port_ms_per_step": 1e3 * float(np.mean(port_t))}
        if extra_sweep:
            line["sweep"] = extra_sweep
        line.update(extra)
        _emit(json.dumps(line))
    p.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweep-batch", type=int, default=64)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
