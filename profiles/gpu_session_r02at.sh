#!/usr/bin/env bash
# round 2, GPU call at: size of the zero-fill launch on the side stream of a fused batch
set -u
cd "$(dirname "$0")/.."
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold --no-e2e"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['ms_per_step_blocks'])"; }
for z in 9472 296 148 592 1184 296 9472; do echo "-- D3H_ZERO_CTAS=$z"; D3H_ZERO_CTAS=$z timeout 300 python bench.py $B 2>/dev/null | show; done
