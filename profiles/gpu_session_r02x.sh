#!/usr/bin/env bash
# round 2, GPU call x: persistent pipelined stream kernel over the transposed rows, timed alone and in the step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02x
run() { env "$@" timeout 200 python profiles/scan_ab.py 2>&1 | tail -1; }
run D3H_SCAN_PIPE=1
run D3H_SCAN_PIPE=1 D3H_SCAN_CPW=1
run D3H_SCAN_PIPE=0
run D3H_SCAN_ROWS=0
echo "== parity (extraction files, all edge paths)"
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
B="--steps 100 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-e2e"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}; print('bench', d['ms_per_step'], d['ms_per_step_blocks'], 'single', d['single_call']['ms_per_frame'], 'cold', (d.get('cold') or {}).get('ms_per_frame'), 'roofline', r['frac'], r['us_per_launch'], 'warm', r['warm_l2_us_per_launch'], 'dev', r.get('device_timer', {}).get('us_per_launch'))"; }
echo "-- pipe"
timeout 400 python bench.py $B 2>gpurun_out/${T}_pipe.err | tee gpurun_out/${T}_pipe.json | show
echo "-- pipe cpw 1"
D3H_SCAN_CPW=1 timeout 400 python bench.py $B 2>gpurun_out/${T}_pipe1.err | tee gpurun_out/${T}_pipe1.json | show
echo "-- csr walk"
D3H_SCAN_ROWS=0 timeout 400 python bench.py $B 2>gpurun_out/${T}_csr.err | tee gpurun_out/${T}_csr.json | show
echo "== device trace, one lane"
timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
echo "== ncu: the pipelined kernel"
D3H_DISABLE_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'edge_scan_pipe_kernel' -s 4 -c 2 -o gpurun_out/${T}_scanpipe python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/${T}_ncu1.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
