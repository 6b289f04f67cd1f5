#!/usr/bin/env bash
# round 2, GPU call h: A/B of the staged stream, the rows-based marking kernel and programmatic dependent launch
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02h
echo "== parity, new defaults (staged scan, rows mark, no PDL)"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -2 gpurun_out/${T}_pytest.log
echo "== parity subset with D3H_PDL=1"
D3H_PDL=1 timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_z_configs.py tests/test_y_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest_pdl.log 2>&1
tail -2 gpurun_out/${T}_pytest_pdl.log
run() { echo "-- $*"; env "$@" timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -2; env "$@" timeout 120 python profiles/single_call_timing.py | tail -1; }
run D3H_SCAN_STAGED=0 D3H_MARK_ROWS=0
run D3H_SCAN_STAGED=1 D3H_MARK_ROWS=0
run D3H_SCAN_STAGED=1 D3H_MARK_ROWS=0 D3H_SCAN_VPT=2
run D3H_SCAN_STAGED=0 D3H_MARK_ROWS=1
run D3H_SCAN_STAGED=1 D3H_MARK_ROWS=1
run D3H_SCAN_STAGED=1 D3H_MARK_ROWS=1 D3H_PDL=1
run D3H_SCAN_STAGED=1 D3H_MARK_ROWS=1 D3H_PDL=1 D3H_SCAN_VPT=2
echo "== 16 frames on 8 lanes"
for e in "D3H_SCAN_STAGED=0 D3H_MARK_ROWS=0" "D3H_SCAN_STAGED=1 D3H_MARK_ROWS=1" "D3H_SCAN_STAGED=1 D3H_MARK_ROWS=1 D3H_SCAN_VPT=2" "D3H_SCAN_STAGED=1 D3H_MARK_ROWS=1 D3H_PDL=1"; do
  echo "-- $e"; env $e timeout 120 python profiles/graph_trace.py --frames 16 --lanes 8 | grep "^#" | grep -v "per frame"
  env $e timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-e2e --no-mesh-stage --no-torch-baseline --no-cold --no-split-pair 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['ms_per_step_blocks'], d['single_call']['ms_per_frame'], d['roofline']['frac'], d['roofline']['device_timer']['frac'])"
done
echo "== ncu of the new kernels"
D3H_DISABLE_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_scan_staged_kernel|edge_mark_rows_kernel|scan_emit_kernel|scan_prefix_kernel' -s 8 -c 8 -o gpurun_out/${T}_scan python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_scan.ncu-rep
