#!/usr/bin/env bash
# round 2, GPU call aw: later rounds of a call keep the topology of the first (no scan / expansion / numbering again)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02aw
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x 2>&1 | tail -2
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold --no-e2e"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['config']['groups'], d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'])"; }
for g in 2 1 4 2 1; do timeout 300 python bench.py $B --groups $g 2>/dev/null | show; done
timeout 300 python profiles/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1
GROUPS=1 timeout 300 python profiles/step_timeline.py > gpurun_out/${T}_timeline_g1.txt 2>&1
grep -v "Warn\|warn" gpurun_out/${T}_timeline_g1.txt | sed -n 1,24p | cut -c1-100; tail -1 gpurun_out/${T}_timeline_g1.txt
