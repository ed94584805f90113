"""Prints the kernels of the LAST tile evaluation in an ncu launch list (gpu__time_duration.sum)."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_tile.csv'
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith('==')]))
names = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')) / 1000, r['Grid Size']) for r in rows]
ends = [i for i, n in enumerate(names) if 'conv_last_bwd' in n[0] or ', 9, 3, ' in n[0]]
start = ends[-2] + 1 if len(ends) > 1 else 0
stop = ends[-1] + 1 if ends else len(names)
if '--step' in sys.argv:     # a whole optimizer step: up to and including the last adam kernel
    adams = [i for i, n in enumerate(names) if 'adam_kernel' in n[0]]
    start, stop = adams[-2] + 1, adams[-1] + 1
if '--all' in sys.argv:      # the whole list (e.g. one step bracketed by cudaProfilerStart/Stop)
    start, stop = 0, len(names)
total, agg = 0.0, OrderedDict()
for name, us, grid in names[start:stop]:
    short = re.sub(r'\(.*', '', name).replace('void ', '').replace('st::', '').replace('<unnamed>::', '')[:64]
    if '-v' in sys.argv:
        print('%-66s %8.1f us  grid %s' % (short, us, grid))
    total += us
    agg.setdefault(short, [0, 0.0])
    agg[short][0] += 1
    agg[short][1] += us
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-66s x%-3d %8.1f us  %5.1f%%' % (k, n, us, 100 * us / total))
print('total %.1f us' % total)
