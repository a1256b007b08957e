"""Selected raw metrics of one ncu capture.  usage: python scripts/ncu_metrics.py gpurun_out/<tag>.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__warps_active.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__waves_per_multiprocessor', 'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__icc_request_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for v in rows[2:]:
    for w in want:
        if w in h: print(w, '=', v[h.index(w)], rows[1][h.index(w)])
    for i, n in enumerate(h):
        if 'issue_stalled' in n and n.endswith('per_issue_active.ratio') and float((v[i] or '0').replace(',', '')) > 0.3: print(n, '=', v[i])
