"""Summarise an `ncu --csv --page raw` log: one line per kernel launch with the metrics that matter here."""
import csv
import sys

WANT = [('Kernel Name', 'kernel', 34), ('gpu__time_duration.sum', 'ns', 8), ('launch__grid_size', 'grid', 6),
        ('launch__registers_per_thread', 'regs', 4), ('dram__bytes_read.sum', 'rdMB', 7), ('dram__bytes_write.sum', 'wrMB', 7),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 6),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%', 6),
        ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1%', 6),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 6),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 6),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%', 6),
        ('sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'tc%bf16', 7),
        ('sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'tc%fp16', 7)]
STALLS = 'smsp__average_warps_issue_stalled_'


def main(path, stalls=False):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    cols = [(hdr.index(k), n, w) for k, n, w in WANT if k in hdr]
    print(' '.join(f'{n:>{w}}' for _, n, w in cols))
    for r in rows[hi + 2:]:
        if len(r) != len(hdr):
            continue
        out = []
        for i, n, w in cols:
            v = r[i]
            if n == 'kernel':
                v = v.replace('void ', '').replace('mc::', '').replace('unnamed>::', '')[:w]
            elif n in ('rdMB', 'wrMB'):
                v = f'{float(v.replace(",", "")) / 1e6:.1f}'
            out.append(f'{v:>{w}}')
        print(' '.join(out))
        if stalls:
            st = sorted(((float(r[i].replace(',', '')), h[len(STALLS):].replace('_per_issue_active.ratio', '')) for i, h in enumerate(hdr)
                         if h.startswith(STALLS) and h.endswith('per_issue_active.ratio') and 'not_issued' not in h and r[i] not in ('', 'n/a')),
                        reverse=True)[:5]
            print('      stalls/issue: ' + ', '.join(f'{n}={v:.2f}' for v, n in st))


if __name__ == '__main__':
    main(sys.argv[1], len(sys.argv) > 2)
