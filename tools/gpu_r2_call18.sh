#!/bin/bash
# Round 2, call 18 (1 GPU): Jacobian snapping + gain kept along a rollout (FD-Riccati provider): parity tests, the suites
# around them, C5 at full size and C1.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call18.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== whitening / snapping tests"
timeout 900 python -m pytest tests/test_gpu_whiten.py -q --timeout 600 2>&1 | tail -25
echo "== neighbours"
timeout 1200 python -m pytest tests/test_gpu_wave.py tests/test_gpu_nearest.py tests/test_gpu_planner.py tests/test_gpu_pipeline.py tests/test_gpu_callbacks.py tests/test_gpu_steer.py -x -q --timeout 900 2>&1 | tail -6
one() { tag=$1; shift
  timeout 600 python bench.py --no-cpu --plan-leg 0 "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -c 600 gpurun_out/bench_r2_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$tag.json')); print('$tag', round(d['value']/1e6,3),'M nodes/s', round(d['ms_per_step'],1),'ms/plan e2e', round(d['e2e']['value']/1e6,3), 'waves', d['config']['waves_per_step'], 'yield', round(d['config']['yield'],3), {k:round(v/d['steps'],1) for k,v in d['stage_ms'].items()})" || true
}
echo "== C5 (FD-Riccati, 1M nodes), C1"
one c5_memo --config c5 --steps 3 --warmup 1
one c1_memo --config c1 --steps 4 --warmup 1
date
one c4_hiprio_sampler --steps 8 --warmup 2
one c5k_hiprio_sampler --config c5k --steps 3 --warmup 1
date
