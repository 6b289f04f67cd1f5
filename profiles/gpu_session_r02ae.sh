#!/usr/bin/env bash
# round 2, GPU call ae: fused frames (grid.y), eligibility fixed, __grid_constant__ frame offsets
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02ae
run() { env "$@" timeout 200 python profiles/scan_ab.py 2>&1 | tail -1; }
run D3H_SCAN_RUNS=1
echo "== parity (extraction files, all edge paths)"
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}; print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'], 'cold', (d.get('cold') or {}).get('ms_per_frame'), 'roofline', r['frac'], r['us_per_launch'], 'warm', r['warm_l2_us_per_launch'], 'dev', r.get('device_timer', {}).get('us_per_launch'), 'e2e', e.get('ms_per_step'), e.get('pos'), 'trace', d.get('device_trace'))"; }
echo "-- fused frames"
timeout 400 python bench.py $B --e2e-pos mapped --e2e-chunk 8 2>gpurun_out/${T}_runs.err | tee gpurun_out/${T}_runs.json | show
tail -3 gpurun_out/${T}_runs.err
echo "-- fused frames, groups 1 / 2"
timeout 400 python bench.py $B --no-e2e --groups 1 2>/dev/null | show
timeout 400 python bench.py $B --no-e2e --groups 2 2>/dev/null | show
echo "-- per-lane graphs"
D3H_FUSE_FRAMES=0 timeout 400 python bench.py $B --no-e2e 2>/dev/null | show
echo "== device trace, one lane / 8 lanes"
timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
timeout 120 python profiles/graph_trace.py --frames 32 --lanes 8 | tail -3
echo "== ncu: scan_runs"
D3H_DISABLE_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'scan_runs_kernel|runs_expand_kernel|poly_cut|poly_faces' -s 12 -c 8 -o gpurun_out/${T}_scanruns python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/${T}_ncu1.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
