#!/bin/bash
# Round 2, call 4: TMA-staged tables + PDL + more old-scan splits: tests, A/B, rollout kernel alone + ncu.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call4.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pipeline + wave tests"
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_wave.py tests/test_gpu_steer.py -q --timeout 600 2>&1 | tail -12
B="timeout 200 python bench.py --steps 8 --warmup 2 --no-cpu --plan-leg 0"
ab() { echo "-- $1"; shift; env "$@" $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'waves',d['config']['waves_per_step'],'digest',d.get('tree_digest'))"; }
ab "device (defaults: PDL on, old scan 32 CTAs/SM)" LQRRT_X=1
ab "device, PDL off" LQRRT_PDL=0
ab "device, tables from global memory" LQRRT_STAGE=0
ab "device, no speculation" LQRRT_SPEC=0
ab "device, lanes 1" LQRRT_TUNING=1,0
ab "host schedule" LQRRT_SCHEDULE=host
echo "== rollout kernel alone"
timeout 200 python tools/k2_once.py
echo "== ncu: rollout kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:steer_kernel -c 2 -o gpurun_out/k2_r2b python tools/k2_once.py > gpurun_out/k2_ncu.log 2>&1; tail -2 gpurun_out/k2_ncu.log
echo "== goal-reaching leg + full bench line"
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 2>&1 | tail -1 > gpurun_out/bench_r2_call4.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_call4.json')); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','plan','cpu_baseline','tree_digest')}, indent=1)); print(json.dumps(d['roofline'])[:600])"
date
