#!/bin/bash
# Round-end evidence run (one gpurun call): GPU tests, smoke, bench (both arms), the ncu launch list
# of the bench command, and an ncu --set full capture of every conv_tc2 launch of one step (DRAM
# traffic per launch for bench.py's roofline.traffic).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "bench ref rc=$?"; cat gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench list rc=$?"
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc2_kernel \
  -o gpurun_out/prof_conv_step_final -f python tools/step_eval.py --steps 3 --profile-last > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30
