#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_scale_n8_final.json 2> gpurun_out/r02_scale_n8_final.err; echo rc=$?
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_scale_n8_final.json').read().strip().splitlines()[-1])
print('N=8: it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()}, 'conv TF/s', round(d['roofline']['achieved']))
for rec in d.get('records', []):
    print('  record:', rec.get('record','')[:40], rec.get('value'), rec.get('unavailable'))
PY
