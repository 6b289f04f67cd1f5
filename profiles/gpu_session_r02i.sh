#!/usr/bin/env bash
# round 2, GPU call i: first run of the tensor-core MLP stage (tests/test_zzzzzz_mlp.py), under compute-sanitizer first
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02i
echo "== MLP building blocks"
timeout 300 python -m pytest tests/test_zzzzzz_mlp.py -m gpu -q -x -k "linear or wgrad or embedding" > gpurun_out/${T}_pytest_blocks.log 2>&1
tail -25 gpurun_out/${T}_pytest_blocks.log | cut -c1-220
echo "== MLP module"
timeout 600 python -m pytest tests/test_zzzzzz_mlp.py -m gpu -q -k "not (linear or wgrad or embedding)" > gpurun_out/${T}_pytest_module.log 2>&1
tail -25 gpurun_out/${T}_pytest_module.log | cut -c1-300
echo "== timing"
timeout 300 python profiles/mlp_bench.py 2>&1 | tail -8
