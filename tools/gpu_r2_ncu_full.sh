#!/bin/bash
# one --set full capture of ONE step of the bench command (profile range), exported to CSV
mkdir -p gpurun_out
ST_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r02_step_full python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i gpurun_out/r02_step_full.ncu-rep --page raw --csv > gpurun_out/r02_step_full_raw.csv 2>/dev/null
ls -la gpurun_out/r02_step_full*; wc -l gpurun_out/r02_step_full_raw.csv
