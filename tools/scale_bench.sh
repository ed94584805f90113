#!/bin/bash
# 1 -> 8 GPU scaling of bench.py on one box (run under gpurun --gpus 8)
mkdir -p gpurun_out
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/scale_$n.json'))
    print('N=$n: %.1f it/s  %.3f ms/step  e2e %.1f  image %.3f ms conv %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown']['image']['ms_per_step'], d['breakdown']['conv_tc']['ms_per_step']))
except Exception as e:
    print('N=$n failed', e); print(open('gpurun_out/scale_$n.err').read()[-1500:])
PY
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_1.err | tail -1 > gpurun_out/scale_1.json
python -c "
import json; d=json.load(open('gpurun_out/scale_1.json')); print('N=1: %.1f it/s  %.3f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
