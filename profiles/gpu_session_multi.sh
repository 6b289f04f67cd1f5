#!/usr/bin/env bash
# multi-GPU session (gpurun --gpus N): real-NCCL parity test, then bench.py in the three modes, each under its own timeout
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
T=${1:-r02m}_n${N}
echo "== $N GPUs: tests/test_multi_gpu.py"
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
launch() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" bench.py --gpus "$N" "${@:2}"; }
echo "== weak (default command of the driver)"
launch 29531 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_weak.json 2> gpurun_out/${T}_weak.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_weak.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'ms_per_step_blocks')}); print('ranks', d['ranks']); e = d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['h2d_only_ms_per_step'], e['h2d_GBps'])
PY
tail -c 300 gpurun_out/${T}_weak.err
echo "== strong (configs[3]: 16 frames per step over all ranks)"
launch 29532 --mode strong --frames-total 16 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_strong.json 2> gpurun_out/${T}_strong.err
python -c "import json; d=json.loads(open('gpurun_out/${T}_strong.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}); print('ranks', d['ranks'])"
tail -c 300 gpurun_out/${T}_strong.err
echo "== tets (configs[4]: 256^3, tet ranges + all-gather of records)"
launch 29533 --mode tets --steps 20 --warmup 3 > gpurun_out/${T}_tets.json 2> gpurun_out/${T}_tets.err
python -c "import json; d=json.loads(open('gpurun_out/${T}_tets.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling','ranks_hold_same_mesh','comm_nranks_ok','ms_per_step_blocks')})"
tail -c 300 gpurun_out/${T}_tets.err
if [ "$N" -ge 4 ]; then
echo "== reference arm under torchrun (rank 0 only)"
launch 29534 --impl reference --steps 5 --warmup 3 | cut -c1-200
fi
