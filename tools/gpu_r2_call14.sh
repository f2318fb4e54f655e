#!/bin/bash
# Round 2, call 14 (1 GPU): fast reset / cached goal validation, sweep of the scan's grid target on whole C4 plans.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call14.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== tests that drive reset / replanning / plans"
timeout 1200 python -m pytest tests/test_gpu_planner.py tests/test_gpu_pipeline.py tests/test_gpu_recognize.py tests/test_gpu_zz_options.py tests/test_gpu_grid.py -x -q --timeout 900 2>&1 | tail -5
one() { tag=$1; shift
  timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu --plan-leg 0 "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -c 400 gpurun_out/bench_r2_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$tag.json')); print('$tag', round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,2), d['e2e']['host_ms_per_step'], 'waves', d['config']['waves_per_step'], 'digest', d['tree_digest'], {k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()})" || true
}
echo "== scan grid target (CTAs per SM) x wave cap"
for w in 16384 8192; do for c in 4 8 12 16 24; do LQRRT_SCAN_CTAS=$c one scan_w${w}_c$c --wave $w; done; done
echo "== wave cap 32768 at 8 CTAs/SM"
LQRRT_SCAN_CTAS=8 one scan_w32768_c8 --wave 32768
date
