#!/usr/bin/env bash
# round 2, GPU call y: run-length compressed edge list (edge_scan_runs_kernel) against the rows and the CSR walk
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02y
run() { env "$@" timeout 200 python profiles/scan_ab.py 2>&1 | tail -1; }
run D3H_SCAN_RUNS=1
run D3H_SCAN_RUNS=0
echo "== parity (extraction files, all edge paths)"
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}; print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'], 'cold', (d.get('cold') or {}).get('ms_per_frame'), 'roofline', r['frac'], r['us_per_launch'], 'warm', r['warm_l2_us_per_launch'], 'dev', r.get('device_timer', {}).get('us_per_launch'), 'e2e', e.get('ms_per_step'), e.get('pos'), 'trace', d.get('device_trace'), 'ranks', d.get('ranks'))"; }
echo "-- runs"
timeout 400 python bench.py $B --e2e-pos mapped --e2e-chunk 8 2>gpurun_out/${T}_runs.err | tee gpurun_out/${T}_runs.json | show
echo "-- rows"
D3H_SCAN_RUNS=0 timeout 400 python bench.py $B --no-e2e 2>gpurun_out/${T}_rows.err | tee gpurun_out/${T}_rows.json | show
echo "-- runs, 16 lanes? (lanes 8 groups 2 / 8)"
timeout 400 python bench.py $B --no-e2e --groups 2 2>/dev/null | show
timeout 400 python bench.py $B --no-e2e --groups 8 2>/dev/null | show
echo "== device trace, one lane"
timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
echo "== ncu: scan + mark"
D3H_DISABLE_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'edge_scan_runs_kernel|edge_mark_kernel' -s 8 -c 4 -o gpurun_out/${T}_scanruns python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/${T}_ncu1.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
