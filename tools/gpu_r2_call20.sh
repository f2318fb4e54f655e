#!/bin/bash
# Round 2, call 20 (8 GPUs, short): the final build at 8 GPUs -- C4 (default command) and C5 (FD-Riccati, whitened search).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call20.log) 2>&1
date
run() { tag=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29680 bench.py --gpus 8 --no-cpu --plan-leg 0 "$@" 2>&1 | tail -1 > gpurun_out/scale_r2_$tag.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_r2_$tag.json'))
    print('$tag N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'host',d['e2e'].get('wave_loop_host_ms_per_step'),'waves',d['config']['waves_per_step'],'digest',d.get('tree_digest'))
except Exception as e:
    print('$tag unreadable', e, open('gpurun_out/scale_r2_$tag.json').read()[-600:])
PY
}
run final_c4_n8 --steps 10 --warmup 3
run final_c5_n8 --config c5 --steps 3 --warmup 1
date
