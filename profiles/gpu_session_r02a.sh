#!/usr/bin/env bash
# round 2, first GPU call: the whole gpu suite WITHOUT -x (every failure visible), smoke, default bench, pair timings
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== parity, all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r02a_pytest.log 2>&1
tail -40 gpurun_out/r02a_pytest.log
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (defaults)"
timeout 400 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 600 gpurun_out/r02a_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02a_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['path_roofline'], d['e2e'], d['single_call'])
print('torch on this GPU:', d.get('torch_gpu_baseline'))
print('mesh stage:', d.get('mesh_stage'))
print('kernels:', d.get('kernels'))
PY
echo "== fused cloth/body pair vs plain split vs two calls (us per pair)"
timeout 200 python - <<'PY'
import time, numpy as np, torch
from d3human_code_b200 import grids
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
dev = torch.device("cuda:0")
pos, tets = grids.kuhn_grid(128)
sdf, msdf = grids.capsule_garment_field(pos)
tp = torch.tensor(pos, device=dev, requires_grad=True); ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
tm = torch.tensor(msdf, device=dev, requires_grad=True); tt = torch.tensor(tets, device=dev)
hm = hmSDF_Tets()
def two():
    return hm(tp, ts, tm, tt, "cloth"), hm(tp, ts, tm, tt, "body")
variants = {"two calls": two, "split": lambda: hm.split(tp, ts, tm, tt), "split fused": lambda: hm.split(tp, ts, tm, tt, fused=True)}
for name, fn in variants.items():
    for bwd in (False, True):
        for _ in range(10):
            c, b = fn()
            if bwd:
                tp.grad = ts.grad = tm.grad = None
                (c[0].sum() + b[0].sum()).backward()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200):
            c, b = fn()
            if bwd:
                tp.grad = ts.grad = tm.grad = None
                (c[0].sum() + b[0].sum()).backward()
        torch.cuda.synchronize()
        print(f"{name:12s} {'fwd+bwd' if bwd else 'fwd    '} {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us")
PY
echo "== device trace, 4 frames on 1 lane, rank table off/on"
for v in 0 1; do D3H_TET_EDGE_RANKS=$v timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -3; done
