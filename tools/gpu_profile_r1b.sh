#!/bin/bash
# One gpurun call: descriptor-offset experiment + ncu --set full of the laggard kernels of one step.
mkdir -p gpurun_out
timeout 60 tools/experiments/desc_offset > gpurun_out/desc_offset.log 2>&1; echo "desc_offset rc=$?"; cat gpurun_out/desc_offset.log | tail -60
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:conv_tc2_kernel<(64|128|16)|conv_first|gram_tc|gram_delta|diff_stats|diff_inject|regularizers|pool_bwd|delta_pack|adam' \
  -o gpurun_out/prof_laggards -f python tools/step_eval.py --steps 3 --profile-last > gpurun_out/ncu_laggards.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_laggards.log
ls -la gpurun_out
