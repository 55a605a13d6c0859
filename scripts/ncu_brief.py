"""Brief per-kernel summary of an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
want = [('time_us', 'gpu__time_duration.sum'), ('sm%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), ('l1%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('l2%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), ('warps%', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'), ('regs', 'launch__registers_per_thread'),
        ('st_long', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
        ('st_short', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
        ('st_lg', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio'),
        ('st_mio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio'),
        ('st_bar', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
        ('st_wait', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),
        ('st_math', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio'),
        ('rdMB', 'dram__bytes_read.sum'), ('wrMB', 'dram__bytes_write.sum'), ('bankconf', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
        ('tc%', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'), ('inst', 'smsp__inst_executed.sum')]
ki = h.index('Kernel Name')
for r in rows[2:]:
    name = r[ki].split('(')[0].split('::')[-1][:44]
    out = [name]
    for label, key in want:
        if key in h:
            v = r[h.index(key)]
            try:
                out.append(f'{label}={float(v.replace(",", "")):.4g}')
            except ValueError:
                out.append(f'{label}={v}')
    print(' '.join(out))
