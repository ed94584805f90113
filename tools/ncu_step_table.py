#!/usr/bin/env python
"""Markdown table of an `ncu --set full` raw CSV export (one row per launch, in launch order)."""
import csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
h, u, d = rows[0], rows[1], rows[2:]
col = {x: i for i, x in enumerate(h)}
def val(r, k, sc=1.0):
    try:
        return float(r[col[k]].replace(',', '')) * sc
    except (KeyError, ValueError):
        return float('nan')
def unit_scale(k):
    return {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u[col[k]], 1)
TENS = ['sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed']
print('| # | kernel | grid | us | DRAM MB (r+w) | DRAM GB/s | DRAM % of peak | tcgen05.mma ops % of peak (UTCHMMA fp16 + bf16) | regs | smem KB |')
print('|---|---|---|---|---|---|---|---|---|---|')
tot = 0.0
for i, r in enumerate(d):
    name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('st::', '').replace('<unnamed>::', '')
    t = val(r, 'gpu__time_duration.sum', unit_scale('gpu__time_duration.sum'))
    byt = val(r, 'dram__bytes_read.sum', unit_scale('dram__bytes_read.sum')) + val(r, 'dram__bytes_write.sum', unit_scale('dram__bytes_write.sum'))
    tot += t
    print('| %d | `%s` | %s | %.1f | %.1f | %.0f | %.0f | %.0f | %d | %.0f |' % (
        i, name[:60], r[col['Grid Size']], t, byt / 1e6, byt / t / 1e3 if t else 0,
        val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), sum(v for v in (val(r, t) for t in TENS) if v == v),
        val(r, 'launch__registers_per_thread'),
        val(r, 'launch__shared_mem_per_block_dynamic', 1e-3 * unit_scale('launch__shared_mem_per_block_dynamic') if 'launch__shared_mem_per_block_dynamic' in col else 1)))
print('\nserialised total %.1f us over %d launches' % (tot, len(d)))
