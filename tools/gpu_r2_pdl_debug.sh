#!/bin/bash
# tight timeouts: ST_PDL_ALL=1 with the early trigger restricted to single-wave grids
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; env "$@" timeout 100 python bench.py --steps 20 --warmup 5 --no-extra --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
t = sys.stdin.read().strip()
if not t: print('NO OUTPUT (hang / crash)'); sys.exit()
d = json.loads(t.splitlines()[-1]); print('it/s %.2f ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})"; }
run default A=1
run pdl_all ST_PDL_ALL=1
run default2 A=1
run pdl_all2 ST_PDL_ALL=1
