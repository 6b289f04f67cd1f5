#!/usr/bin/env bash
# round 2, GPU call w: stream kernel variants timed alone (rows phased CPW 1/2/4, unphased, CSR walk); e2e with mapped positions by chunk size
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02w
run() { env "$@" timeout 200 python profiles/scan_ab.py 2>&1 | tail -1; }
run D3H_SCAN_CPW=2
run D3H_SCAN_CPW=1
run D3H_SCAN_CPW=4
run D3H_SCAN_PHASED=0
run D3H_SCAN_ROWS=0
run D3H_SCAN_CPW=2
B="--steps 60 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-cold"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}; print('bench', d['ms_per_step'], 'single', d['single_call']['ms_per_frame'], 'roofline', r['frac'], r['us_per_launch'], 'warm', r['warm_l2_us_per_launch'], 'dev', r.get('device_timer', {}).get('us_per_launch'), 'e2e', e.get('ms_per_step'), e.get('value'), e.get('pos'), e.get('chunk_frames'))"; }
for c in 4 8 16; do
  echo "-- mapped e2e, chunk $c"
  timeout 400 python bench.py $B --e2e-pos mapped --e2e-chunk $c 2>gpurun_out/${T}_mapped$c.err | tee gpurun_out/${T}_mapped$c.json | show
done
echo "-- copy e2e, chunk 8"
timeout 400 python bench.py $B --e2e-chunk 8 2>gpurun_out/${T}_copy8.err | tee gpurun_out/${T}_copy8.json | show
echo "== device trace, one lane"
timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
echo "== ncu: the rows kernel"
D3H_DISABLE_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'edge_scan_rows_kernel' -s 4 -c 2 -o gpurun_out/${T}_scanrows python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/${T}_ncu1.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
