#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02n
timeout 300 python profiles/mlp_blocks_timing.py 2>&1 | tee gpurun_out/${T}_blocks.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_wgrad_kernel' -s 3 -c 2 -o gpurun_out/${T}_wgrad python profiles/mlp_blocks_timing.py > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_wgrad.ncu-rep
