#!/bin/bash
# GPU tests + the five BASELINE configurations on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tools/config_runs.py > gpurun_out/config_runs.log 2>&1; echo "configs rc=$?"; grep -E "^cfg" gpurun_out/config_runs.log; tail -3 gpurun_out/config_runs.log | cut -c1-600
