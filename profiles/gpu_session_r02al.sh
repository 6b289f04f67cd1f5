#!/usr/bin/env bash
# round 2, GPU call al: where the e2e step goes (positions resident / results kept on the device), prepare of the shared path
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02al
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "fused_frames or mapped or batch or frames" 2>&1 | tail -2
B="--steps 60 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold --e2e-pos mapped"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}; print('bench', d['ms_per_step'], 'e2e', e.get('ms_per_step'), e.get('value'), e.get('pos'), e.get('chunk_frames'))"; }
echo "-- mapped"; timeout 400 python bench.py $B 2>/dev/null | show
echo "-- mapped, results stay on the device"; D3H_E2E_DIAG=nod2h timeout 400 python bench.py $B 2>/dev/null | show
echo "-- positions resident"; D3H_E2E_DIAG=resident timeout 400 python bench.py $B 2>/dev/null | show
echo "-- positions resident, results stay"; D3H_E2E_DIAG=resident,nod2h timeout 400 python bench.py $B 2>/dev/null | show
timeout 300 python profiles/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1
grep -v "Warn\|warn" gpurun_out/${T}_timeline.txt | sed -n 1,12p | cut -c1-110; tail -1 gpurun_out/${T}_timeline.txt
