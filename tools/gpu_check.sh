#!/bin/bash
# quick GPU iteration: parity check of the tensor-core kernels, launch list of one tile evaluation
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py check 2>&1 | grep -v "^features" | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tile.csv python tools/tile_eval.py --evals 2 2>&1 | tail -2
timeout 300 python tools/tc_check.py time 2>&1 | tail -2
