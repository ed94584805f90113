#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc32_check.py 2>&1 | grep -v "features" | tail -12
export ST_PARITY_REPORT=$PWD/gpurun_out/parity_tc32.jsonl; rm -f $ST_PARITY_REPORT
timeout 600 python -m pytest tests/test_gpu_parity_configs.py tests/test_gpu_parity.py -m gpu -q --tb=line -k "tc32 and not cfg3" 2>&1 | tail -6
python - <<'PY'
import json
for l in open('gpurun_out/parity_tc32.jsonl'):
    d = json.loads(l); print(d['case'], {k: (round(v, 7) if isinstance(v, float) else v) for k, v in d.items() if k in ('loss_rel', 'grad_l2rel', 'frac_gt_1e3', 'rms', 'max')})
PY
for env in "A=1"; do
env $env timeout 200 python bench.py --precision tc32 --no-cpu-baseline --no-e2e --no-extra --steps 6 --warmup 2 2>/dev/null | python -c "
import sys, json
t = sys.stdin.read().strip()
if not t: print('NO OUTPUT'); sys.exit()
d = json.loads(t.splitlines()[-1]); print('$env tc32 it/s %.2f ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})"
done
