#!/usr/bin/env bash
# round 2, GPU call g: full suite, the new bench (modes, blocks, cold, split pair, compact e2e), reference arm, configs[0]/[4],
# compute-sanitizer on hardware, ncu launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02g
echo "== parity (all gpu tests)"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (defaults, every leg)"
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'ms_per_step_blocks')})
print('roofline', d['roofline']['frac'], d['roofline'].get('device_timer', {}).get('frac'), 'path', d['path_roofline']['frac_step'], d['path_roofline']['frac_single_call'])
e = d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e['h2d_only_ms_per_step'], e['h2d_GBps'])
print('single', d['single_call']); print('cold', d['cold']); print('split_pair', d['split_pair']); print('ranks', d['ranks'])
print('kernels:', {k: v['us_avg'] for k, v in d.get('kernels', {}).items()})
print('trace', d['device_trace'])
print('cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None, 'torch gpu', d['torch_gpu_baseline'])
PY
echo "== reference arm"
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
cut -c1-300 gpurun_out/${T}_bench_ref.json
echo "== frames API instead of packed"
timeout 300 python bench.py --api frames --steps 100 --no-cpu-baseline --no-e2e --no-mesh-stage --no-torch-baseline --no-cold --no-split-pair 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_blocks'], d['ranks'])"
echo "== configs[0]: 64^3 sphere"
timeout 300 python bench.py --res 64 --field sphere --steps 100 --no-mesh-stage --no-torch-baseline --no-split-pair > gpurun_out/${T}_bench_c0.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/${T}_bench_c0.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['single_call'], d['cold'], d['cpu_baseline']['value'])"
echo "== configs[4]: 256^3 tet-sharded path on one GPU"
timeout 400 python bench.py --mode tets --steps 20 > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
cut -c1-600 gpurun_out/${T}_bench_c4.json; tail -c 300 gpurun_out/${T}_bench_c4.err
echo "== compute-sanitizer memcheck / racecheck (24^3, every path)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitizer_case.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitizer_case.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${T}_racecheck.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --blocks 1 --no-e2e --no-cpu-baseline --no-torch-baseline --no-mesh-stage --no-cold --no-split-pair > gpurun_out/${T}_launches.log 2>&1
wc -l gpurun_out/${T}_launches.csv
