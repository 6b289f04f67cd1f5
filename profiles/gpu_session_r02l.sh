#!/usr/bin/env bash
# round 2, GPU call j: MLP building-block tests again (tolerance), ncu of the tensor-core kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02l
timeout 300 python -m pytest tests/test_zzzzzz_mlp.py -m gpu -q > gpurun_out/${T}_pytest_mlp.log 2>&1
tail -5 gpurun_out/${T}_pytest_mlp.log | cut -c1-220
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_wgrad_kernel|wgrad_reduce_kernel|head_backward_kernel|mlp_pack' -s 2 -c 8 -o gpurun_out/${T}_mlp python profiles/mlp_bench.py --points 200000 --reps 1 > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_mlp.ncu-rep
timeout 300 python profiles/mlp_bench.py 2>&1 | tail -1 > gpurun_out/r02l_mlp_bench.json; cat gpurun_out/r02l_mlp_bench.json | cut -c1-900
