#!/usr/bin/env bash
# round 2, GPU call ao: the A/B switches still pass the parity files (per-lane graphs, topology per frame, no run-length tables)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
F="tests/test_cuda_parity.py tests/test_z_configs.py tests/test_zzzz_fused_pair.py"
echo "-- D3H_FUSE_FRAMES=0"; D3H_FUSE_FRAMES=0 timeout 600 python -m pytest $F -m gpu -q -x 2>&1 | tail -1
echo "-- D3H_SHARE_TOPOLOGY=0"; D3H_SHARE_TOPOLOGY=0 timeout 600 python -m pytest $F -m gpu -q -x 2>&1 | tail -1
echo "-- D3H_SCAN_RUNS=0"; D3H_SCAN_RUNS=0 timeout 600 python -m pytest $F -m gpu -q -x 2>&1 | tail -1
echo "-- D3H_SCAN_PIPE=1 D3H_SCAN_RUNS=0"; D3H_SCAN_PIPE=1 D3H_SCAN_RUNS=0 timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "golden or oracle" 2>&1 | tail -1
echo "-- bench smoke with 5 / 12 / 40 frames per rank (rounds, tails)"
for f in 5 12 40; do timeout 300 python bench.py --frames-per-rank $f --steps 20 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['frames_per_step'], d['ms_per_step'], d['value'])"; done
