#!/bin/bash
# 2-GPU box: multi-GPU test, the GPU suite (small cases), bench at N=1 (PDL on/off) and N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -s 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q --tb=short -x 2>&1 | tail -15
run_bench() {  # name, extra env, args
  name=$1; shift
  env "$@" timeout 900 python bench.py --no-cpu-baseline --steps 20 --warmup 5 $BENCH_ARGS > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$name.json'))
    print('$name: it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %s (%.2f of value) launches %d' % (d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e'] and round(d['e2e']['value'],1), d['e2e'] and d['e2e']['fraction_of_value'] or 0, d['gpu_launches']))
    print('  roofline frac %.3f (%s) ' % (d['roofline']['frac'], d['roofline']['peak_kind']), {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
    for r in d.get('records', []):
        print('  record:', r.get('record','')[:40], r.get('value'), r.get('ms_per_step'), r.get('unavailable'))
except Exception as e:
    print('bench parse failed', e)
PY
}
BENCH_ARGS="" run_bench n1_pdl A=1
BENCH_ARGS="--no-extra --no-e2e" run_bench n1_nopdl ST_NO_PDL=1
for pdl in 1 0; do
  if [ $pdl = 0 ]; then export ST_NO_PDL=1; else unset ST_NO_PDL; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/bench_n2_pdl$pdl.json 2> gpurun_out/bench_n2_pdl$pdl.err; tail -3 gpurun_out/bench_n2_pdl$pdl.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n2_pdl$pdl.json').read().strip().splitlines()[-1])
    print('n2 pdl=$pdl: it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %s launches %d' % (d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e'] and round(d['e2e']['value'],1), d['gpu_launches']))
    print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
done
