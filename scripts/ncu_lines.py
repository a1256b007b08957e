"""Per-source-line totals of one ncu --set full --import-source on capture: executed warp instructions and stall samples.
usage: python scripts/ncu_lines.py gpurun_out/<tag>.ncu-rep [instr|samples] [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
by = sys.argv[2] if len(sys.argv) > 2 else "instr"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg, text, cur, key = collections.defaultdict(lambda: [0, 0, 0]), {}, None, None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        h = r; si, ie = h.index('# Samples'), h.index('Instructions Executed'); continue
    if r[0].isdigit():
        key = (cur, int(r[0])); text[key] = r[1]
    elif len(r) > 2 and r[2].startswith('0x') and key:
        agg[key][0] += int(r[si] or 0); agg[key][1] += int(r[ie] or 0); agg[key][2] += 1
ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("total executed warp instructions %d, samples %d; sorted by %s" % (ti, ts, by))
acc = 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1 if by == "instr" else 0])[:top]:
    acc += v[1 if by == "instr" else 0]
    print('instr %5.1f%%  samples %5.1f%%  cum %5.1f%%  static %4d  %s:%d  %s' % (100.0 * v[1] / ti, 100.0 * v[0] / ts, 100.0 * acc / (ti if by == "instr" else ts), v[2], k[0], k[1], text[k].strip()[:100]))
