#!/usr/bin/env bash
# round 2, GPU call ay: final state (topology kept across the rounds of a call)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02ay
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
r = d['roofline']; print('roofline', r['kernel'], r['frac'], r['us_per_launch'], r['algorithmic_bytes_per_launch'], r.get('share_of_frame_kernel_time'), r.get('traffic'))
print('all', {k: (v['us_per_launch'], round(v['frac'], 3)) for k, v in r['all_kernels'].items()})
print('path', d['path_roofline']['frac_step'], d['path_roofline']['frac_single_call'])
e = d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e['pos'], 'copy:', (e.get('positions_copied_first') or {}).get('ms_per_step'))
print('single', d['single_call']['ms_per_frame'], 'cold', d['cold']['ms_per_frame'], 'split', {k: v for k, v in d['split_pair'].items() if k.endswith('_ms')})
print('cpu', d['cpu_baseline']['value'], 'torch gpu', d['torch_gpu_baseline']['ms_per_frame'], 'clocks', d.get('clocks'))
PY
echo "== reference arm"
timeout 400 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
cut -c1-200 gpurun_out/${T}_bench_ref.json
NGROUPS=1 timeout 300 python profiles/step_timeline.py > gpurun_out/${T}_timeline_g1.txt 2>&1
grep -v "Warn\|warn" gpurun_out/${T}_timeline_g1.txt | sed -n 1,22p | cut -c1-100; tail -1 gpurun_out/${T}_timeline_g1.txt
