"""Selected raw metrics + top stall reasons of an ncu report: ncu -i x.ncu-rep --page raw --csv | python scripts/ncu_rep_summary.py <title>"""
import csv
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
print('# ' + ' '.join(sys.argv[1:]))
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    for k in KEEP:
        if k in hdr:
            print(f'{k:100s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}')
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('per_issue_active.ratio') and r[i] not in ('', 'n/a')]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f'{h:100s} {v}')
