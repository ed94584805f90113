#!/bin/bash
# one --set full capture of ONE step of the bench command (profile range); only the CSV export and a
# per-kernel summary travel back (the .ncu-rep is ~90 MB: over the 64 MiB limit of gpurun_out/)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
ST_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none -f -o /tmp/r02_step_full python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-120
ncu -i /tmp/r02_step_full.ncu-rep --page raw --csv > gpurun_out/r02_step_full_raw.csv 2>/dev/null
wc -l gpurun_out/r02_step_full_raw.csv; du -sh gpurun_out
