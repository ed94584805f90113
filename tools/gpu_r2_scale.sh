#!/bin/bash
# 8-GPU box: bench at N = 8, 4, 2, 1 (no extras), then the two-GPU bit-identity test
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2 1; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-extra > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -2 gpurun_out/scale_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1])
    print('N=$n: it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %s launches %d' % (d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e'] and round(d['e2e']['value'],1), d['gpu_launches']))
    print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()}, 'conv TF/s', round(d['roofline']['achieved']))
except Exception as e:
    print('bench parse failed', e)
PY
done
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -3
