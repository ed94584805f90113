#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc32_check.py > gpurun_out/tc32_check.log 2>&1; tail -40 gpurun_out/tc32_check.log
rm -f gpurun_out/parity_report.jsonl
export ST_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q --tb=line -k "tc32 and not cfg3 and not cfg4" 2>&1 | tail -30 > gpurun_out/pytest_configs.log
tail -12 gpurun_out/pytest_configs.log
cat gpurun_out/parity_report.jsonl
for p in tc32; do
timeout 600 python bench.py --precision $p --no-cpu-baseline --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err; tail -2 gpurun_out/bench_$p.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$p.json'))
    print('$p: it/s %.2f  ms/step %.3f  launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))
    print({k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
done
