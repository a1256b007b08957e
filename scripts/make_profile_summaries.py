"""Turns gpurun_out/{launches_*.csv, prof_*.ncu-rep, bench_*.json} into the committed summaries under profiles/.
usage: python scripts/make_profile_summaries.py r1c r1"""
import collections, csv, io, os, shutil, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "%s_launches.csv" % out))
shutil.copy(os.path.join(G, "bench_%s.json" % tag), os.path.join(P, "bench_%s.json" % out))
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
idx = [hdr.index(k) for k in keep if k in hdr]
with open(os.path.join(P, "%s_ncu_full_summary.csv" % out), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:reach_build"], capture_output=True, text=True).stdout
agg, text, cur = collections.defaultdict(lambda: [0, 0, 0]), {}, None
stall = collections.Counter()
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        h = r; si, ie, sb = h.index('# Samples'), h.index('Instructions Executed'), h.index('stall_barrier')
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not' not in c]
        continue
    if r[0].isdigit():
        key = (cur, int(r[0])); text[key] = r[1]
        try:
            agg[key][0] += int(r[si] or 0); agg[key][1] += int(r[ie] or 0); agg[key][2] += int(r[sb] or 0)
            for i, c in stall_cols: stall[c] += int(r[i] or 0)
        except (ValueError, IndexError):
            pass
ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
with open(os.path.join(P, "%s_reach_hot_lines.txt" % out), "w") as f:
    f.write("reach_build_kernel: warp-stall samples per source line (ncu --set full, source page)\n")
    tot = sum(stall.values())
    f.write("stall reasons: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in stall.most_common(8)) + "\n")
    f.write("samples% instr% barrier-share  file:line  source\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        f.write('%5.1f %5.1f %5.1f  %s:%d  %s\n' % (100.0 * v[0] / ts, 100.0 * v[1] / ti, 100.0 * v[2] / max(1, v[0]), k[0], k[1], text[k].strip()[:100]))
print(open(os.path.join(P, "%s_reach_hot_lines.txt" % out)).read()[:1500])
