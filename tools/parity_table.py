#!/usr/bin/env python
"""gpurun_out/parity_report.jsonl (written by tests/test_gpu_parity_configs.py under
ST_PARITY_REPORT) -> a markdown table for profiles/."""
import json, sys

rows = [json.loads(l) for l in open(sys.argv[1])]
print('| case | mode | targets | loss rel | grad rel L2 | grad max rel | pixels > 1e-3 max | q50 | q99 | q99.9 |')
print('|---|---|---|---|---|---|---|---|---|---|')
for r in rows:
    if 'grad_l2rel' not in r:
        continue
    print('| %s | %s | %s | %.1e | %.2e | %.2e | %.2e | %.1e | %.1e | %.1e |' % (
        r['case'], r['precision'], r.get('targets', 'oracle'), r['loss_rel'], r['grad_l2rel'],
        r['grad_maxrel'], r.get('frac_gt_1e3', float('nan')), r.get('q50', float('nan')),
        r.get('q99', float('nan')), r.get('q999', float('nan'))))
print()
print('| case (image after N iterations, grey levels of 0..255) | mode | max | RMS | pixels > 1 | q99 |')
print('|---|---|---|---|---|---|')
for r in rows:
    if 'rms' not in r:
        continue
    print('| %s | %s | %.2f | %.3f | %s | %s |' % (
        r['case'], r['precision'], r['max'], r['rms'],
        '%.2e' % r['frac_gt1'] if 'frac_gt1' in r else '-', '%.2f' % r['q99'] if 'q99' in r else '-'))
