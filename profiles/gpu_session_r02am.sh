#!/usr/bin/env bash
# round 2, GPU call am: mapped e2e leg with every chunk launched up front
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02am
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "fused_frames or mapped or batch or frames" 2>&1 | tail -2
B="--steps 60 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold --e2e-pos mapped"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}; print('bench', d['ms_per_step'], 'e2e', e.get('ms_per_step'), e.get('value'), e.get('pos'), e.get('chunk_frames'))"; }
echo "-- mapped"; timeout 400 python bench.py $B 2>/dev/null | show
echo "-- mapped, results stay on the device"; D3H_E2E_DIAG=nod2h timeout 400 python bench.py $B 2>/dev/null | show
echo "-- positions resident"; D3H_E2E_DIAG=resident timeout 400 python bench.py $B 2>/dev/null | show
echo "-- positions resident, results stay"; D3H_E2E_DIAG=resident,nod2h timeout 400 python bench.py $B 2>/dev/null | show
echo "-- mapped, chunk 16 / 4"; timeout 400 python bench.py $B --e2e-chunk-mapped 16 2>/dev/null | show; timeout 400 python bench.py $B --e2e-chunk-mapped 4 2>/dev/null | show
