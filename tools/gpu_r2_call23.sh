#!/bin/bash
# Round 2, call 23 (4 GPUs, short): the missing point of the final build's scaling table.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call23.log) 2>&1
date
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --plan-leg 0 2>&1 | tail -1 > gpurun_out/scale_r2_final_c4_n4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/scale_r2_final_c4_n4.json'))
print('N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'waves',d['run']['waves_per_step'],'digest',d.get('tree_digest'))
PY
date
