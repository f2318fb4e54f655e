#!/bin/bash
# Round 2, call 3: fixed tests, A/B of rollout lanes / CTA size / speculation on whole plans, ncu of the rollout kernel.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call3.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pipeline + wave + recognise + planner tests"
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_wave.py tests/test_gpu_recognize.py tests/test_gpu_planner.py tests/test_gpu_nearest.py -q --timeout 600 2>&1 | tail -25
echo "== A/B on whole plans (C4, 8 plans each)"
B="timeout 200 python bench.py --steps 8 --warmup 2 --no-cpu --plan-leg 0"
ab() { echo "-- $1"; shift; env "$@" $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'waves',d['config']['waves_per_step'],'digest',d.get('tree_digest'))"; }
ab "device, lanes 8 (default)" LQRRT_X=1
ab "device, lanes 1" LQRRT_TUNING=1,0
ab "device, lanes 8, steer block 32" LQRRT_STEER_BLOCK=32
ab "device, lanes 1, steer block 32" LQRRT_TUNING=1,0 LQRRT_STEER_BLOCK=32
ab "device, no speculation, lanes 8" LQRRT_SPEC=0
ab "device, no speculation, lanes 1" LQRRT_SPEC=0 LQRRT_TUNING=1,0
ab "host schedule" LQRRT_SCHEDULE=host
ab "device, scan 8 CTAs/SM" LQRRT_SCAN_CTAS=8
ab "device, scan 32 CTAs/SM" LQRRT_SCAN_CTAS=32
echo "== rollout kernel alone"
timeout 200 python tools/k2_once.py
echo "== ncu: rollout kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:steer_kernel -c 4 -o gpurun_out/k2_r2 python tools/k2_once.py > gpurun_out/k2_ncu.log 2>&1; tail -3 gpurun_out/k2_ncu.log
date
