#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tile1024.csv python tools/tile_eval.py --evals 2 --size 1024 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 36 -c 24 -o gpurun_out/prof_conv_tc2 -f python tools/tile_eval.py --evals 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
