#!/bin/bash
# Round 2, call 22 (2 GPUs): the driver's own multi-GPU command, unabridged (goal-reaching leg and CPU baseline included),
# and the reference arm under torchrun.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call22.log) 2>&1
date
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/driver_n2.json 2> gpurun_out/driver_n2.err; echo "rc=$?"; tail -c 600 gpurun_out/driver_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/driver_n2.json').read().strip().splitlines()[-1])
print('N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2), 'digest', d['tree_digest'], 'clocks', d['clocks'])
print('plan', d.get('plan')); print('cpu', d.get('cpu_baseline'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
date
