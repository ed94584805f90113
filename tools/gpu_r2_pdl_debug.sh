#!/bin/bash
# tight timeouts: which launches hang with ST_PDL_ALL=1?
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; env "$@" timeout 90 python bench.py --steps 10 --warmup 3 --no-extra --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-160; echo "rc=$?"; }
run default A=1
run pdl_all ST_PDL_ALL=1
run pdl_all_1tile ST_PDL_ALL=1 ST_MAX_BATCH=1
