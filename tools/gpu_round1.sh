#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list + full capture of the conv kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tile.csv python tools/tile_eval.py --evals 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 36 -c 12 -o gpurun_out/prof_conv_tc python tools/tile_eval.py --evals 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
