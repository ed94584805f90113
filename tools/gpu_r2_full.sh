#!/bin/bash
# full GPU suite with the parity report, then bench lines of the fp16 and tc32 modes
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
export ST_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
timeout 2400 python -m pytest tests -m gpu -q --tb=short -rs --durations=12 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
unset ST_PARITY_REPORT
for p in tc32 fp16; do
timeout 600 python bench.py --precision $p --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err; tail -2 gpurun_out/bench_$p.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$p.json'))
    print('$p: it/s %.2f  ms/step %.3f  e2e %s launches %d' % (d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['value'], d['gpu_launches']))
    print({k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
done
