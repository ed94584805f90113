#!/bin/bash
# One gpurun call per kernel iteration: parity tests first (stop on failure), then bench, the ncu
# launch list of one step and an ncu --set full capture of the small-layer conv kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_gpu.log; exit $rc; fi
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_latest.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_latest.json').read().strip().splitlines()[-1])
    print('it/s %.2f  ms/step %.3f  e2e %.2f  launches %d  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'share_of_step')})
    print({k: round(v['ms_per_step'], 3) for k, v in d['breakdown'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_step.csv python tools/step_eval.py --steps 3 --profile-last > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/launch_table.py gpurun_out/launches_step.csv --all -v 2>/dev/null | head -120
timeout 120 python tools/step_eval.py --steps 3 2>&1 | tail -2
timeout 120 python tools/step_eval.py --steps 3 --size 1024 2>&1 | tail -2
if [ "${FULL:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base mangled \
  -k 'regex:conv_tc2_kernelILi(64|128|16)ELi9' -o gpurun_out/prof_conv_small -f \
  python tools/step_eval.py --steps 3 --profile-last > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
fi
ls -la gpurun_out | head -30
