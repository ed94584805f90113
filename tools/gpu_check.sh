#!/bin/bash
# quick GPU iteration: GPU tests, parity of the tensor-core kernels, launch list of one evaluation
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/tc_check.py check 2>&1 | grep -v "^features" | tail -8
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_latest.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_latest.json'))
    print('it/s %.2f  ms/step %.3f  e2e %.2f  launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
    print('roofline', d['roofline'])
    print({k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
timeout 300 python bench.py --no-cpu-baseline --no-e2e --precision fp16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fp16 mode: it/s %.2f  conv frac %.3f' % (d['value'], d['roofline']['frac']))"
