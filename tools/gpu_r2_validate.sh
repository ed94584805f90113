#!/bin/bash
# end-of-round validation: what the driver runs (GPU suite, smoke, both bench arms at N = 1)
mkdir -p gpurun_out
export ST_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl; rm -f $ST_PARITY_REPORT
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rs 2>&1 | tail -12
unset ST_PARITY_REPORT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
bash tools/gpu_r2_final.sh
