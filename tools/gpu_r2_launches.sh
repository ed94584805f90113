#!/bin/bash
mkdir -p gpurun_out
ST_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
python tools/launch_table.py gpurun_out/r02_launches_raw.csv --all -v > gpurun_out/r02_launches.md 2>&1; tail -90 gpurun_out/r02_launches.md
