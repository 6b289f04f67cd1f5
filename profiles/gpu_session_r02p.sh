#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02p
timeout 300 python -m pytest tests/test_zzzzzz_mlp.py -m gpu -q > gpurun_out/${T}_pytest_mlp.log 2>&1
tail -5 gpurun_out/${T}_pytest_mlp.log | cut -c1-220
timeout 300 python profiles/mlp_bench.py 2>&1 | tail -1 > gpurun_out/${T}_mlp_bench.json; cat gpurun_out/${T}_mlp_bench.json | cut -c1-1000
