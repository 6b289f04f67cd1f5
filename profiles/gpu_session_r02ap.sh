#!/usr/bin/env bash
# round 2, GPU call ap: register caps of the two surface kernels (variants built with -DD3H_FACES_MINB / -DD3H_CUT_MINB)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold --no-e2e"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'])"; }
cp d3human-code_b200/lib/libd3h_tets.so /tmp/keep.so
for v in f3c4 f4c4 f4c5 f5c5 f3c4 f4c4; do
  cp d3human-code_b200/lib/variants/$v.so d3human-code_b200/lib/libd3h_tets.so
  echo "-- $v"; timeout 300 python bench.py $B 2>/dev/null | show
done
cp /tmp/keep.so d3human-code_b200/lib/libd3h_tets.so
