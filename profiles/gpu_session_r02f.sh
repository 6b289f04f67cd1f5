#!/usr/bin/env bash
# round 2, GPU call f: sub-queues, barrier-free scan (VPT variants)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity (all gpu tests, three edge paths)"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_pytest.log 2>&1
tail -3 gpurun_out/r02f_pytest.log
for v in 4 2 1; do
echo "== device trace, 4 frames on 1 lane, D3H_SCAN_VPT=$v"
D3H_SCAN_VPT=$v timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -2
done
echo "-- sphere field"; timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 --field sphere | tail -2
echo "== device trace, 16 frames on 8 lanes"
timeout 120 python profiles/graph_trace.py --frames 16 --lanes 8 | grep "^#" | grep -v "per frame"
echo "== bench"
timeout 400 python bench.py --no-cpu-baseline --no-torch-baseline --no-mesh-stage > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -c 300 gpurun_out/r02f_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['path_roofline'], d['e2e']['value'] if d.get('e2e') else None, d['single_call'])
print('kernels:', {k: v['us_avg'] for k, v in d.get('kernels', {}).items()})
PY
echo "== ncu"
D3H_DISABLE_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_scan_kernel|edge_mark_kernel|poly_cut_kernel|poly_faces_kernel|scan_emit_kernel|scan_prefix_kernel|prepare_kernel|adjoint' -s 20 -c 10 -o gpurun_out/r02f_scan python profiles/graph_trace.py --frames 2 --lanes 1 > gpurun_out/r02f_ncu.log 2>&1
ls -la gpurun_out/r02f_scan.ncu-rep
