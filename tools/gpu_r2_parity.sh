#!/bin/bash
# round 2: parity at the benchmark's shapes (tests/test_gpu_parity_configs.py) with the measurement
# report, the un-skipped jitter test and the set_params golden; then the rest of the GPU suite
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
export ST_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q --tb=short -rs --durations=15 2>&1 | tail -60 > gpurun_out/pytest_configs.log
tail -5 gpurun_out/pytest_configs.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q --tb=short --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
