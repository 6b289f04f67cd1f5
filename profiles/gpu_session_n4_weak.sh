#!/usr/bin/env bash
# weak-scaling point at the number of GPUs of the box (default command of the driver)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
T=${1:-r02av}_n${N}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus "$N" --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_weak.json 2> gpurun_out/${T}_weak.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_weak.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'ms_per_step_blocks')}); print('ranks', d['ranks']['local_step_ms'], d['ranks']['allreduce_ms']); e = d['e2e']; print('e2e', e['value'], e['ms_per_step'])
PY
