#!/usr/bin/env bash
# round 2, GPU call aj: e2e leg through the packed API; new fused / regrowth test
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02aj
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "fused_frames or mapped or batch" 2>&1 | tail -2
B="--steps 60 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}; c=e.get('positions_copied_first') or {}; print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'], 'e2e', e.get('ms_per_step'), e.get('value'), e.get('pos'), e.get('chunk_frames'), e.get('h2d_bytes_per_step'), e.get('d2h_bytes_per_step'), 'copy', c.get('ms_per_step'))"; }
timeout 400 python bench.py $B 2>gpurun_out/${T}_a.err | tee gpurun_out/${T}_a.json | show
tail -2 gpurun_out/${T}_a.err
timeout 400 python bench.py $B --e2e-pos mapped --e2e-chunk-mapped 16 2>/dev/null | show
timeout 400 python bench.py $B --e2e-pos mapped --e2e-chunk-mapped 4 2>/dev/null | show
