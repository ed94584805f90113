#!/bin/bash
# quick iteration: the small GPU suite + bench N=1 (no extras) + config runner for cfg2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q --tb=short -x 2>&1 | tail -15
timeout 900 python bench.py --no-cpu-baseline --no-extra --steps 30 --warmup 5 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_iter.json'))
    print('it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %s (%.2f of value) launches %d  clocks %s' % (d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e'] and round(d['e2e']['value'],1), d['e2e'] and d['e2e']['fraction_of_value'] or 0, d['gpu_launches'], d['clocks']))
    print('  roofline frac %.3f (%s) ' % (d['roofline']['frac'], d['roofline']['peak_kind']), {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
timeout 600 python tools/config_runs.py cfg2 cfg4 2>&1 | tail -3 | cut -c1-600
