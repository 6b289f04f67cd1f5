#!/usr/bin/env bash
# round 2, GPU call au: programmatic dependent launch once more, now that the kernels of the chain are 5-10 us
set -u
cd "$(dirname "$0")/.."
B="--steps 60 --no-cpu-baseline --no-mesh-stage --no-torch-baseline --no-split-pair --no-sdf-query --no-lbs-stage --no-e2e"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['single_call']; c=d['cold']; print('bench', d['ms_per_step'], 'single', s['ms_per_frame'], s['fwd_ms_median'], s['bwd_ms_median'], 'cold', c['ms_per_frame'], c['fwd_ms_median'])"; }
for p in 0 1 0 1; do echo "-- D3H_PDL=$p"; D3H_PDL=$p timeout 300 python bench.py $B 2>/dev/null | show; done
echo "-- trace, one lane"; timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
D3H_PDL=1 timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
