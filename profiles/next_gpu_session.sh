#!/usr/bin/env bash
# First GPU call of the next session: everything that was written after round 1's GPU budget ran out, in order of value,
# each step under its own timeout so that a hang cannot hold the box.  Usage (1 GPU):
#   gpurun --timeout 900 -- 'bash profiles/next_gpu_session.sh 2>&1 | tee gpurun_out/next_session.log'
# then (2 GPUs, charged twice):  gpurun --gpus 2 --timeout 400 -- 'bash profiles/next_gpu_session.sh multi'
# and the ncu evidence:          gpurun --timeout 1500 -- 'bash profiles/next_gpu_session.sh profile'
set -u
cd "$(dirname "$0")/.."
if [ "${1:-}" = "multi" ]; then
  N=$(python -c "import torch; print(torch.cuda.device_count())")
  timeout 150 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus "$N" --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
  exit 0
fi
if [ "${1:-}" = "profile" ]; then
  # ncu evidence for the CURRENT defaults (32 frames / 4 groups / 8 lanes): launch list + one --set full capture of the
  # dominant kernel and of the mesh kernels.  Never a bench value: ncu serialises and replays.  Copy the results from
  # gpurun_out/ to profiles/r02_* and summarise with launch_summary.py / ncu -i ... --page raw --csv.
  mkdir -p gpurun_out
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-torch-baseline > gpurun_out/r02_launches.log 2>&1
  tail -1 gpurun_out/r02_launches.log | cut -c1-200
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:classify_kernel -c 3 -o gpurun_out/r02_classify \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-torch-baseline --no-mesh-stage > gpurun_out/r02_classify.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mesh_ -c 24 -o gpurun_out/r02_mesh \
      python profiles/mesh_bench.py --reps 3 > gpurun_out/r02_mesh.log 2>&1
  ls -la gpurun_out | tail -8
  exit 0
fi
echo "== parity (incl. the tests never run on a GPU: test_z_configs.py, test_zy_config5.py, test_zz_mesh.py, test_zzzz_fused_pair.py)"
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (defaults)"
timeout 300 python bench.py 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['roofline'].get('device_timer', {}).get('frac'),
      d['path_roofline']['frac_step'], d['e2e']['value'] if d['e2e'] else None, d['single_call'])
print('torch on this GPU:', d.get('torch_gpu_baseline'))
print('mesh stage:', d.get('mesh_stage'))"
echo "== fused cloth/body pair vs plain split vs two calls (us per pair, fwd only and fwd+bwd)"
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
echo "== mesh stage (8f row 2): Mesh.edges / auto_normals vs the same torch ops"
timeout 200 python profiles/mesh_bench.py 2>&1 | tail -1
echo "== per-tet edge-rank table variant (compact_kernel<3>) vs bisection: single-lane trace and bench"
for v in 0 1; do
  echo "-- D3H_TET_EDGE_RANKS=$v"
  D3H_TET_EDGE_RANKS=$v timeout 120 python profiles/graph_trace.py --frames 4 --lanes 1 | tail -1
  D3H_TET_EDGE_RANKS=$v timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9)"
done
echo "== device trace, 16 frames on 8 lanes"
timeout 120 python profiles/graph_trace.py --frames 16 --lanes 8 | grep "^#" | grep -v "per frame"
