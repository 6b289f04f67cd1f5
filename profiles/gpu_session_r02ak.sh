#!/usr/bin/env bash
# round 2, GPU call ak: shared records / corner ids, full-grid prepare of the first frame
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02ak
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x 2>&1 | tail -2
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold --no-e2e"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'])"; }
timeout 400 python bench.py $B 2>gpurun_out/${T}_a.err | tee gpurun_out/${T}_a.json | show
tail -2 gpurun_out/${T}_a.err
timeout 300 python profiles/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1
grep -v "Warn\|warn" gpurun_out/${T}_timeline.txt | sed -n 1,14p | cut -c1-110; tail -1 gpurun_out/${T}_timeline.txt
