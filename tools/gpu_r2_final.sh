#!/bin/bash
# the two arms exactly as the driver runs them at N = 1
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -2 gpurun_out/r02_bench_reference.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
python - <<PY
import json
r = json.load(open('gpurun_out/r02_bench_reference.json'))
print('reference: %.5f it/s  ms/step %.0f  full_iteration %s  cores %s' % (r['value'], r['ms_per_step'], r['full_iteration'] and round(r['full_iteration']['seconds'],1), r['cpu_baseline']['cores']))
d = json.load(open('gpurun_out/r02_bench_n1.json'))
print('engine: it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %.1f (%.2f of value) launches %d clocks %s' % (d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['e2e']['fraction_of_value'], d['gpu_launches'], d['clocks']))
print('  roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','peak_kind','traffic','share_of_step')})
print('  gram', d['roofline_gram'] and {k: d['roofline_gram'][k] for k in ('achieved','frac')})
print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
for rec in d.get('records', []):
    print('  record:', rec.get('record','')[:50], rec.get('value'), rec.get('ms_per_step'), rec.get('unavailable'), rec.get('roofline') and round(rec['roofline']['frac'],3))
print('  cpu', d['cpu_baseline'] and {k: d['cpu_baseline'][k] for k in ('value','cores')}, d['cpu_baseline'] and d['cpu_baseline'].get('best_case',{}).get('value'))
PY
