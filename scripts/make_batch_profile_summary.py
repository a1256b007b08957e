"""Turns one `ncu --set full --import-source on` capture of the sweep-shaped reach_build_kernel (gpurun_out/<tag>.ncu-rep,
made with `ncu ... python scripts/tune_batch.py one 128 16 0`) into profiles/<out>_batch_ncu_summary.csv and
profiles/<out>_batch_hot_lines.txt.   usage: python scripts/make_batch_profile_summary.py batch2_full r1"""
import collections, csv, io, os, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", tag + ".ncu-rep")
P = os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__icc_request_hit_rate.pct', 'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
        'gcc__average_cache_request_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
idx = [hdr.index(k) for k in keep if k in hdr]
with open(os.path.join(P, "%s_batch_ncu_summary.csv" % out), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg, text, cur = collections.defaultdict(lambda: [0, 0, 0]), {}, None
stall = collections.Counter()
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        h = r; si, ie = h.index('# Samples'), h.index('Instructions Executed')
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not' not in c]
        continue
    if r[0].isdigit():
        key = (cur, int(r[0])); text[key] = r[1]
    elif len(r) > 2 and r[2].startswith('0x'):
        agg[key][0] += int(r[si] or 0); agg[key][1] += int(r[ie] or 0); agg[key][2] += 1
        for i, c in stall_cols: stall[c] += int(r[i] or 0)
ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
with open(os.path.join(P, "%s_batch_hot_lines.txt" % out), "w") as f:
    f.write("reach_build_kernel (" + tag + "): warp-stall samples and executed warp instructions per source line (ncu --set full, source page)\n")
    tot = sum(stall.values())
    f.write("stall reasons: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in stall.most_common(8)) + "\n")
    f.write("static SASS instructions %d, executed warp instructions %d\n" % (sum(v[2] for v in agg.values()), ti))
    f.write("samples% instr% static  file:line  source\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        f.write('%5.1f %5.1f %5d  %s:%d  %s\n' % (100.0 * v[0] / ts, 100.0 * v[1] / ti, v[2], k[0], k[1], text[k].strip()[:100]))
print(open(os.path.join(P, "%s_batch_hot_lines.txt" % out)).read()[:1200])
