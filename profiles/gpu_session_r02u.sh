#!/usr/bin/env bash
# round 2, GPU call r: the state the driver will see -- full gpu suite, smoke, default bench, reference arm
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02u
echo "== parity (all gpu tests)"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (defaults, every leg)"
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'ms_per_step_blocks')})
print('roofline', d['roofline']['frac'], d['roofline'].get('device_timer', {}).get('frac'), 'path', d['path_roofline']['frac_step'], d['path_roofline']['frac_single_call'])
e = d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e['h2d_only_ms_per_step'])
print('single', d['single_call']['ms_per_frame'], 'cold', d['cold']['ms_per_frame'], 'split', {k: v for k, v in d['split_pair'].items() if k.endswith('_ms')})
print('sdf_query', d['sdf_query'])
print('mesh', {k: (v.get('edges_us'), v.get('normals_fwd_bwd_us')) for k, v in d['mesh_stage'].items() if isinstance(v, dict)})
print('cpu', d['cpu_baseline']['value'], 'torch gpu', d['torch_gpu_baseline']['ms_per_frame'])
PY
echo "== reference arm"
timeout 400 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
cut -c1-200 gpurun_out/${T}_bench_ref.json
python - <<PY
import json
d = json.loads(open('gpurun_out/r02u_bench.json').read().strip().splitlines()[-1])
print('roofline full', d['roofline'])
print('lbs', d['lbs_stage'])
PY
echo "== memcheck on the tensor-core stage (small cases)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_zzzzzz_mlp.py -m gpu -q -k "golden or (linear and 77) or (wgrad and 33) or head_backward and 515" 2>&1 | tail -3
