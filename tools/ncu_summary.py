#!/usr/bin/env python3
"""Summarise an ncu report exported with
   ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
Usage: python tools/ncu_summary.py raw.csv src.csv"""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print('%-70s %-12s %s' % (w, units[i], ' '.join(r[i] for r in data)))
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct') and 'not_issued' not in h:
        try:
            v = float(data[0][i])
        except ValueError:
            continue
        if v > 3:
            print('stall %-60s %.1f' % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
rows = list(csv.reader(open(src)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
start = hdr_idx[0]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
h = rows[start]
ci, cs, csamp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
hist, samp = collections.Counter(), collections.Counter()
tot = tots = 0
nwarps = None
for r in rows[start + 1:end]:
    if len(r) <= ci:
        continue
    try:
        n, s = int(r[ci]), int(r[csamp])
    except ValueError:
        continue
    if nwarps is None:
        nwarps = n
    toks = r[cs].strip().split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    hist[op] += n
    samp[op] += s
    tot += n
    tots += s
print('total warp-inst %d  per warp %.0f  samples %d' % (tot, tot / max(nwarps, 1), tots))
for op, n in hist.most_common(18):
    print('%-8s %9d %5.1f%%  per-warp %6.0f  stall-samples %5.1f%%' % (op, n, 100 * n / tot, n / nwarps, 100 * samp[op] / max(tots, 1)))
