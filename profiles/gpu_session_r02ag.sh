#!/usr/bin/env bash
# round 2, GPU call ag: shared topology for the frames of a batch that share sdf / msdf
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02ag
echo "== parity (extraction files, all edge paths)"
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}; print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'], 'cold', (d.get('cold') or {}).get('ms_per_frame'), 'e2e', e.get('ms_per_step'), e.get('pos'))"; }
echo "-- shared topology"
timeout 400 python bench.py $B --e2e-pos mapped --e2e-chunk 8 2>gpurun_out/${T}_shared.err | tee gpurun_out/${T}_shared.json | show
tail -3 gpurun_out/${T}_shared.err
echo "-- groups 2 / 1"
timeout 400 python bench.py $B --no-e2e --groups 2 2>/dev/null | show
timeout 400 python bench.py $B --no-e2e --groups 1 2>/dev/null | show
echo "-- topology per frame"
D3H_SHARE_TOPOLOGY=0 timeout 400 python bench.py $B --no-e2e 2>/dev/null | show
timeout 300 python profiles/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1
grep -v "Warn\|warn" gpurun_out/${T}_timeline.txt | head -28 | cut -c1-120; tail -1 gpurun_out/${T}_timeline.txt
