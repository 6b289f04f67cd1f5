#!/usr/bin/env bash
# round 2, GPU call s: tangent-branch backward on hardware + the full suite + sanitizer on the new kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02s
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -k "tangent" 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
timeout 300 python profiles/single_call_timing.py | tail -1
cat > /tmp/tng_case.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from d3human_code_b200 import grids
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
dev = torch.device("cuda:0")
pos, tets = grids.kuhn_grid(20); sdf, msdf = grids.capsule_garment_field(pos)
tp = torch.tensor(pos, device=dev, requires_grad=True); ts = torch.tensor(sdf, device=dev, requires_grad=True); tm = torch.tensor(msdf, device=dev, requires_grad=True)
for _ in range(2):
    v, f, _, _, t, ex = hmSDF_Tets()(tp, ts, tm, torch.tensor(tets, device=dev), "cloth")
    (t.sum() + ex["v_tng_watertight"].square().sum() + v.sum()).backward()
torch.cuda.synchronize(); print("tangent sanitizer case ok", float(tp.grad.abs().max()))
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/tng_case.py 2>&1 | tail -3
