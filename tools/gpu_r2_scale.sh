#!/bin/bash
# 8-GPU box: bench at N = 8 (default, ST_PDL_ALL=1, and with the extra records as the driver runs it), 4, 2, 1
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('%s: it/s %.2f  ms/step %.3f  host enqueue %.3f ms  e2e %s launches %d' % (sys.argv[2], d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e'] and round(d['e2e']['value'],1), d['gpu_launches']))
    print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()}, 'conv TF/s', round(d['roofline']['achieved']))
    for rec in d.get('records', []):
        print('   record:', rec.get('record','')[:40], rec.get('value'), rec.get('unavailable'))
except Exception as e:
    print(sys.argv[2], 'bench parse failed', e)
PY
}
torun() { n=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((10 + RANDOM % 80)) bench.py --gpus $n --steps 30 --warmup 5 "$@"; }
torun 8 --no-extra > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err; show gpurun_out/r02_scale_n8.json N=8
ST_PDL_ALL=1 torun 8 --no-extra --no-e2e > gpurun_out/r02_scale_n8_pdlall.json 2> gpurun_out/r02_scale_n8_pdlall.err; show gpurun_out/r02_scale_n8_pdlall.json N=8-pdl-all
torun 4 --no-extra > gpurun_out/r02_scale_n4.json 2> gpurun_out/r02_scale_n4.err; show gpurun_out/r02_scale_n4.json N=4
torun 2 --no-extra > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err; show gpurun_out/r02_scale_n2.json N=2
timeout 150 python bench.py --gpus 1 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err; show gpurun_out/r02_scale_n1.json N=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_scale_n8_full.json 2> gpurun_out/r02_scale_n8_full.err; show gpurun_out/r02_scale_n8_full.json N=8-as-driver
