#!/usr/bin/env bash
# round 2, GPU call q: persistent pipelined edge scan A/B, MLP tests, strong mode at N=1, mesh-kernel ncu
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02q
timeout 300 python -m pytest tests/test_zzzzzz_mlp.py -m gpu -q 2>&1 | tail -2
run() { echo "-- $*"; env "$@" timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1; }
run D3H_SCAN_PIPE=0
run D3H_SCAN_PIPE=1 D3H_SCAN_VPT_PIPE=2
run D3H_SCAN_PIPE=1 D3H_SCAN_VPT_PIPE=1
for e in "D3H_SCAN_PIPE=0" "D3H_SCAN_PIPE=1 D3H_SCAN_VPT_PIPE=2" "D3H_SCAN_PIPE=1 D3H_SCAN_VPT_PIPE=1"; do
  echo "-- bench $e"
  env $e timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-e2e --no-mesh-stage --no-torch-baseline --no-cold --no-split-pair --no-sdf-query 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['ms_per_step_blocks'], d['single_call']['ms_per_frame'], 'roofline', d['roofline']['frac'], d['roofline']['device_timer']['frac'], d['roofline']['us_per_launch'])"
done
echo "== parity with the pipelined scan"
D3H_SCAN_PIPE=1 timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x 2>&1 | tail -2
echo "== strong mode (configs[3]: 16 frames) on one GPU"
timeout 300 python bench.py --mode strong --frames-total 16 --steps 100 --no-cpu-baseline --no-e2e --no-mesh-stage --no-torch-baseline --no-cold --no-split-pair --no-sdf-query > gpurun_out/${T}_strong_n1.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/${T}_strong_n1.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['path_roofline']['frac_step'])"
echo "== ncu: pipelined scan + mesh kernels"
D3H_SCAN_PIPE=1 D3H_DISABLE_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'edge_scan_pipe_kernel' -s 4 -c 2 -o gpurun_out/${T}_scanpipe python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/${T}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'mesh_' -c 20 -o gpurun_out/${T}_mesh python profiles/mesh_bench.py --reps 2 > gpurun_out/${T}_ncu2.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
