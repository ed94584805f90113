#!/usr/bin/env python
"""DRAM traffic per launch from an ncu `--set full` capture, for bench.py's `roofline.traffic`.

    ncu -i gpurun_out/conv_step.ncu-rep --page raw --csv > gpurun_out/conv_step_raw.csv
    python tools/ncu_traffic.py gpurun_out/conv_step_raw.csv --size 2048 --tile 512 --precision fp16 --gpus 1

Adds / replaces the entries of profiles/ncu_traffic.json:
  "<kernel>|size=..|tile=..|precision=..|gpus=.." -> {dram_bytes_per_launch, launches, source}
(dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches of each kernel)."""
import argparse, csv, json, os, re

p = argparse.ArgumentParser()
p.add_argument('csv')
p.add_argument('--size', type=int, default=2048)
p.add_argument('--tile', type=int, default=512)
p.add_argument('--precision', default='fp16')
p.add_argument('--gpus', type=int, default=1)
p.add_argument('--out', default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json'))
a = p.parse_args()
rows = list(csv.reader(l for l in open(a.csv) if not l.startswith('==')))
header, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(header)}
def scale(unit):
    return {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
agg = {}
for r in data:
    name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('st::', '').replace('<unnamed>::', '')
    base = re.sub(r'<.*', '', name)
    m = re.match(r'conv_tc2_kernel<\s*\d+,\s*(\d+)', name)
    if m and m.group(1) == '1':
        base = 'conv_tc2_kernel_1x1'          # the style GEMM; 'conv_tc2_kernel' = the 3x3 convolutions
    rd = float(r[col['dram__bytes_read.sum']].replace(',', '')) * scale(units[col['dram__bytes_read.sum']])
    wr = float(r[col['dram__bytes_write.sum']].replace(',', '')) * scale(units[col['dram__bytes_write.sum']])
    t = float(r[col['gpu__time_duration.sum']].replace(',', ''))
    e = agg.setdefault(base, [0, 0.0, 0.0])
    e[0] += 1; e[1] += rd + wr; e[2] += t
table = json.load(open(a.out)) if os.path.exists(a.out) else {}
for base, (n, byt, t) in agg.items():
    key = '%s|size=%d|tile=%d|precision=%s|gpus=%d' % (base, a.size, a.tile, a.precision, a.gpus)
    table[key] = {'dram_bytes_per_launch': byt / n, 'launches': n,
                  'source': 'ncu --set full --clock-control none, %s (dram__bytes_read.sum + dram__bytes_write.sum, mean over %d launches of one step)' % (os.path.basename(a.csv), n)}
    print('%-40s x%-3d %8.1f MB/launch' % (base, n, byt / n / 1e6))
json.dump(table, open(a.out, 'w'), indent=1, sort_keys=True)
