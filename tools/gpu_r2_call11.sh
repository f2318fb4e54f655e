#!/bin/bash
# Round 2, call 11 (1 GPU): first run of the indexed nearest search (csrc/grid.cuh): parity tests, then whole plans with
# the index off / on at several size thresholds and wave caps, the goal-reaching leg, and an ncu launch list.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call11.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== grid tests"
timeout 900 python -m pytest tests/test_gpu_grid.py -x -q --timeout 600 2>&1 | tail -15
echo "== neighbours: nearest / pipeline / options / planner"
timeout 1200 python -m pytest tests/test_gpu_nearest.py tests/test_gpu_pipeline.py tests/test_gpu_zz_options.py tests/test_gpu_planner.py tests/test_gpu_wave.py -x -q --timeout 900 2>&1 | tail -6
b() { tag=$1; shift
  timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu --plan-leg 0 "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -c 400 gpurun_out/bench_r2_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$tag.json')); print('$tag', round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,2), 'waves', d['config']['waves_per_step'], 'digest', d['tree_digest'], {k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()}, d['e2e']['host_ms_per_step'])" || true
}
echo "== C4 whole plans: index off, on (thresholds), wave caps, lag 1"
LQRRT_GRID=0 b grid0
b grid_auto
LQRRT_GRID_MIN=0 b grid_min0
LQRRT_GRID_MIN=4096 b grid_min4k
LQRRT_GRID_MIN=65536 b grid_min64k
b grid_w16k --wave 16384
b grid_w4k --wave 4096
b grid_lag1 --lag 1
LQRRT_GRID=0 b c3_grid0 --config c3
b c3_grid --config c3
LQRRT_GRID=0 b c2_grid0 --config c2 --steps 4
b c2_grid --config c2 --steps 4
echo "== goal-reaching leg"
LQRRT_GRID=0 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --plan-leg 6 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('grid off', d['plan'])"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --plan-leg 6 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('grid on ', d['plan'])"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu --plan-leg 0 > gpurun_out/launches_r2.log 2>&1; tail -1 gpurun_out/launches_r2.log | cut -c1-200
echo "== ncu --set full of the indexed search and the chain kernels (converted to csv here; the reports stay on the box)"
timeout 600 ncu --set full --clock-control none -k regex:'nearest_grid|steer_x|commit_x' -s 90 -c 6 -o /tmp/ncu_grid_r2 -f python bench.py --steps 1 --warmup 0 --no-cpu --plan-leg 0 > gpurun_out/ncu_grid_r2.log 2>&1; tail -2 gpurun_out/ncu_grid_r2.log | cut -c1-200
ncu -i /tmp/ncu_grid_r2.ncu-rep --page raw --csv > gpurun_out/ncu_grid_r2_raw.csv 2>/dev/null; ls -la gpurun_out/ncu_grid_r2_raw.csv
du -sh gpurun_out
date
