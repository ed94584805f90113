#!/bin/bash
# 2-GPU check (gpurun --gpus 2): N-rank result == 1-rank result bit for bit, then bench.py at N=2.
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py 2>&1 | grep -E "rank|Error|error" | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 2> gpurun_out/scale_2.err | tail -1 > gpurun_out/scale_2.json
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/scale_2.json'))
    print('N=2: %.1f it/s  %.3f ms/step  e2e %.1f  image %.3f ms conv %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown']['image']['ms_per_step'], d['breakdown']['conv_tc']['ms_per_step']))
except Exception as e:
    print('N=2 failed', e); print(open('gpurun_out/scale_2.err').read()[-1500:])
PY
