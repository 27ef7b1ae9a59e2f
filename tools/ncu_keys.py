#!/usr/bin/env python
"""Print the metrics we judge kernels by from `ncu -i X.ncu-rep --page raw --csv` output (one column per launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_mio_throttle.ratio',
        'smsp__average_warp_latency_issue_stalled_lg_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio',
        'smsp__average_warp_latency_issue_stalled_not_selected.ratio', 'smsp__average_warp_latency_issue_stalled_wait.ratio',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']
extra = sys.argv[2:]
for w in want + extra:
    hits = [h for h in hdr if h == w] or ([h for h in hdr if w in h] if w in extra else [])
    for h in hits:
        i = hdr.index(h)
        print(f"{h} [{units[i]}]: " + " | ".join(r[i][:60] for r in rows[2:]))
