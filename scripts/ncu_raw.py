"""Dev tool: key raw metrics of the first kernel in an .ncu-rep. usage: python scripts/ncu_raw.py report.ncu-rep"""
import csv, subprocess, sys, io
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in h: print("%-70s %s %s" % (w, v[h.index(w)], u[h.index(w)]))
for i, n in enumerate(h):
    if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio'):
        try:
            x = float(v[i])
        except ValueError:
            continue
        if x > 0.15: print("stall %-40s %.2f" % (n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], x))
