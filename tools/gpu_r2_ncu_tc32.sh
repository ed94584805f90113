#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
ST_NCU_RANGE=1 timeout 700 ncu --profile-from-start off --set full --clock-control none -f -o /tmp/r02_tc32_full python bench.py --precision tc32 --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ncu_tc32.log 2>&1
tail -1 gpurun_out/ncu_tc32.log | cut -c1-120
ncu -i /tmp/r02_tc32_full.ncu-rep --page raw --csv > gpurun_out/r02_tc32_step_full_raw.csv 2>/dev/null
wc -l gpurun_out/r02_tc32_step_full_raw.csv; du -sh gpurun_out
