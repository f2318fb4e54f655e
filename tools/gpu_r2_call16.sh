#!/bin/bash
# Round 2, call 16 (1 GPU): state-independent metric through the whitened diagonal scan (LQ_OPT_WHITEN_METRIC): parity
# tests, the suites around it, C5 at full size with the option on and off.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call16.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== whitening tests"
timeout 900 python -m pytest tests/test_gpu_whiten.py -x -q --timeout 600 2>&1 | tail -15
echo "== neighbours"
timeout 1200 python -m pytest tests/test_gpu_wave.py tests/test_gpu_nearest.py tests/test_gpu_planner.py tests/test_gpu_pipeline.py -x -q --timeout 900 2>&1 | tail -6
one() { tag=$1; shift
  timeout 600 python bench.py --no-cpu --plan-leg 0 "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -c 600 gpurun_out/bench_r2_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$tag.json')); print('$tag', round(d['value']/1e6,3),'M nodes/s', round(d['ms_per_step'],1),'ms/plan e2e', round(d['e2e']['value']/1e6,3), 'waves', d['config']['waves_per_step'], 'yield', round(d['config']['yield'],3), {k:round(v/d['steps'],1) for k,v in d['stage_ms'].items()})" || true
}
echo "== C5 (FD-Riccati, 1M nodes): whitened metric on / off"
one c5_whiten --config c5 --steps 2 --warmup 1
LQRRT_WHITEN=0 one c5_nowhiten --config c5 --steps 1 --warmup 0
echo "== C5 with waves of at most 8192"
one c5_whiten_w8192 --config c5 --steps 2 --warmup 1 --wave 8192
date
