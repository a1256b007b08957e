"""Turns the final round-2 captures under gpurun_out/ into the committed summaries under profiles/:
  r2_final.ncu-rep (ncu --set full, one step: reach_build / hyperplane / constraint_eval) -> r2_ncu_full_summary.csv,
  r2_reach_hot_lines.txt, r2_traffic.json (DRAM bytes per launch + digest of the CUDA sources the capture was made from)
  r2_final_launches.csv -> r2_launches.csv
usage: python scripts/make_r2_profiles.py"""
import csv, hashlib, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, ROOT)


def source_digest():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "armour-dev_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")) and name != "armour_capi.cu":   # device code only: the host-side C ABI file does not change the kernels
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


rep = os.path.join(G, "r2_final.ncu-rep")
if not os.path.exists(rep):
    rep = os.path.join(ROOT, "ncu_raw", "r2_final.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__icc_request_hit_rate.pct',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
idx = [hdr.index(k) for k in keep if k in hdr]
with open(os.path.join(P, "r2_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
traffic = {"source": "profiles/r2_ncu_full_summary.csv (ncu --set full --clock-control none, one launch each, cold caches)", "source_digest": source_digest()}
ir, iw, iu = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), None
units = rows[1]
scale = lambda u: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    name = "reach_build_kernel" if "reach_build" in r[ik] else "constraint_eval_kernel" if "constraint_eval" in r[ik] else "hyperplane_kernel" if "hyperplane" in r[ik] else None
    if name:
        traffic[name] = float(r[ir].replace(",", "")) * scale(units[ir]) + float(r[iw].replace(",", "")) * scale(units[iw])
json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
shutil.copy(os.path.join(G, "r2_final_launches.csv"), os.path.join(P, "r2_launches.csv"))
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, "samples", "40"], capture_output=True, text=True).stdout
open(os.path.join(P, "r2_reach_hot_lines.txt"), "w").write("all three kernels of one step (reach_build_kernel dominates); ncu --set full, source page\n" + out)
print(json.dumps(traffic, indent=1)); print(out[:1500])
