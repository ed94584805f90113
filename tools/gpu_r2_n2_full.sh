#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo rc=$?; tail -3 gpurun_out/r02_bench_n2.err | cut -c1-300
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print('N=2: it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['gpu_launches']))
for rec in d.get('records', []):
    print('  record:', rec.get('record','')[:50], rec.get('value'), rec.get('ms_per_step'), rec.get('unavailable'))
PY
