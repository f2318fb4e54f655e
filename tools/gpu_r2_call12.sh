#!/bin/bash
# Round 2, call 12 (8 GPUs): strong scaling of the pipelined (lag 2) wave driver at 1/2/4/8 GPUs, wave caps 8192 and
# 16384, lag 1 at 8 GPUs for comparison, C5 (FD-Riccati, 1M nodes) at 8 GPUs, and the multi-GPU parity tests.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call12.log) 2>&1
date; nvidia-smi --query-gpu=name --format=csv,noheader | head -8
echo "== grid tests"
timeout 300 python -m pytest tests/test_gpu_grid.py -x -q --timeout 200 2>&1 | tail -3
echo "== multi-GPU tests"
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 2>&1 | tail -4
run() { n=$1; tag=$2; shift; shift
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --plan-leg 0 "$@" 2>&1 | tail -1 > gpurun_out/scale_r2_$tag.json
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --plan-leg 0 "$@" 2>&1 | tail -1 > gpurun_out/scale_r2_$tag.json; fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_r2_$tag.json'))
    print('$tag N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'host',d['e2e'].get('wave_loop_host_ms_per_step'),'waves',d['config']['waves_per_step'],'digest',d.get('tree_digest'))
except Exception as e:
    print('$tag unreadable', e, open('gpurun_out/scale_r2_$tag.json').read()[-600:])
PY
}
for n in 1 2 4 8; do run $n n$n; done
echo "== wave cap 16384"
run 1 w16k_n1 --wave 16384
run 8 w16k_n8 --wave 16384
echo "== lag 1 at 8 GPUs"
run 8 lag1_n8 --lag 1
echo "== C5 (FD-Riccati, 1M nodes) at 8 GPUs"
run 8 c5_n8 --config c5 --steps 2 --warmup 1
echo "== C5k at 8 GPUs"
run 8 c5k_n8 --config c5k --steps 3 --warmup 1
date
