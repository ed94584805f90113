#!/bin/bash
timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
t = sys.stdin.read().strip()
if not t: print('NO OUTPUT'); sys.exit()
d = json.loads(t.splitlines()[-1]); print('it/s %.2f ms/step %.3f e2e %.1f (%.3f)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['fraction_of_value']))"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
