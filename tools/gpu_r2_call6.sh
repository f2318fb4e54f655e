#!/bin/bash
# Round 2, call 6 (8 GPUs): scaling of the device-scheduled pipeline, 1/2/4/8 GPUs, host-side pipeline timers.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call6.log) 2>&1
date; nvidia-smi --query-gpu=name --format=csv,noheader | head -8
echo "== multi-GPU tests (2..4 ranks, peer and nccl)"
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 2>&1 | tail -8
run() { n=$1; shift
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --plan-leg 0 "$@" 2>&1 | tail -1 > gpurun_out/s$n.json
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --plan-leg 0 "$@" 2>&1 | tail -1 > gpurun_out/s$n.json; fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s$n.json'))
    print('N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'host',d['e2e'].get('wave_loop_host_ms_per_step'),'waves',d['config']['waves_per_step'],'digest',d.get('tree_digest'))
except Exception as e:
    print('N=$n unreadable', e, open('gpurun_out/s$n.json').read()[-800:])
PY
}
for n in 1 2 4 8; do run $n; done
echo "== 8 GPUs, goal-reaching leg only check + nccl exchange for comparison"
LQRRT_EXCHANGE=nccl run 8
mv gpurun_out/s8.json gpurun_out/s8_nccl.json
date
