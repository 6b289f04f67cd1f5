#!/usr/bin/env bash
# round 2, GPU call v: transposed edge rows (edge_scan_rows_kernel) A/B against the CSR walk; mapped-position probe of the e2e leg
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02v
echo "== parity (extraction files, all edge paths)"
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}; print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'], 'cold', (d.get('cold') or {}).get('ms_per_frame'), 'roofline', r['frac'], r['us_per_launch'], 'warm', r['warm_l2_us_per_launch'], 'dev', r.get('device_timer', {}).get('us_per_launch'), 'bytes', r['algorithmic_bytes_per_launch'], 'e2e', e.get('ms_per_step'), e.get('value'), e.get('pos'), e.get('h2d_bytes_per_step'))"; }
echo "-- rows"
timeout 400 python bench.py $B 2>gpurun_out/${T}_rows.err | tee gpurun_out/${T}_rows.json | show
echo "-- csr walk"
D3H_SCAN_ROWS=0 timeout 400 python bench.py $B --no-e2e 2>gpurun_out/${T}_csr.err | tee gpurun_out/${T}_csr.json | show
echo "-- rows, mapped positions in the e2e leg"
timeout 400 python bench.py $B --e2e-pos mapped 2>gpurun_out/${T}_mapped.err | tee gpurun_out/${T}_mapped.json | show
tail -5 gpurun_out/${T}_mapped.err
echo "== device trace, one lane"
timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
D3H_SCAN_ROWS=0 timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
echo "== ncu: the rows kernel"
D3H_DISABLE_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'edge_scan_rows_kernel' -s 4 -c 2 -o gpurun_out/${T}_scanrows python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/${T}_ncu1.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
