#!/bin/bash
# Round 2, call 5 (2 GPUs): rollout instruction diet + first multi-GPU run of the device-scheduled pipeline.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call5.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
echo "== callbacks / steer / pipeline / wave tests after the diet"
timeout 900 python -m pytest tests/test_gpu_callbacks.py tests/test_gpu_steer.py tests/test_gpu_pipeline.py tests/test_gpu_wave.py tests/test_gpu_planner.py -q --timeout 600 2>&1 | tail -12
echo "== multi-GPU tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 2>&1 | tail -12
echo "== rollout kernel alone"
timeout 200 python tools/k2_once.py
echo "== bench 1 GPU"
timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu --plan-leg 0 2>&1 | tail -1 > gpurun_out/b1.json
echo "== bench 2 GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 2 --no-cpu --plan-leg 0 2>&1 | tail -1 > gpurun_out/b2.json
python - <<'PY'
import json
for f in ('gpurun_out/b1.json','gpurun_out/b2.json'):
    try:
        d=json.load(open(f))
        print(f, d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'waves',d['config']['waves_per_step'],'digest',d.get('tree_digest'))
    except Exception as e:
        print(f, 'unreadable', e, open(f).read()[-600:])
PY
date
