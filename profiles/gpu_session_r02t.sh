#!/usr/bin/env bash
# round 2, GPU call t: first run of the skinning stage + the rewritten MLP head backward
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02t
timeout 300 python -m pytest tests/test_zzzzzzz_lbs.py -m gpu -q > gpurun_out/${T}_pytest_lbs.log 2>&1; tail -12 gpurun_out/${T}_pytest_lbs.log | cut -c1-250
timeout 300 python -m pytest tests/test_zzzzzz_mlp.py -m gpu -q 2>&1 | tail -3 | cut -c1-250
timeout 300 python profiles/lbs_bench.py 2>&1 | tail -1 | tee gpurun_out/${T}_lbs_bench.json | cut -c1-700
timeout 300 python profiles/mlp_bench.py 2>&1 | tail -1 | tee gpurun_out/${T}_mlp_bench.json | cut -c1-500
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_zzzzzzz_lbs.py -m gpu -q -k "golden or extracted" 2>&1 | tail -3
