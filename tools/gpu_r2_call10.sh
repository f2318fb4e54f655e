#!/bin/bash
# Round 2, call 10 (1 GPU): full GPU suite at HEAD, ncu launch list of the bench command (fixed skip count), ncu --set
# full of the scan at the saturated working point, wave-cap sweep of the throughput leg, full C5 (FD-Riccati, 1M nodes)
# once, ncu of the Riccati rollout / metric kernels.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call10.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -6
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu --plan-leg 0 > gpurun_out/launches_r2.log 2>&1; tail -1 gpurun_out/launches_r2.log | cut -c1-200
grep -c nearest_diag_pk gpurun_out/launches_r2.csv; grep -c nearest_diag_kernel gpurun_out/launches_r2.csv
echo "== ncu --set full: packed and scalar scan at n = 6, W = 8192, N = 1e5"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nearest_diag -c 2 -o gpurun_out/ncu_k1_r2 -f python tools/k1_once.py > gpurun_out/ncu_k1_r2.log 2>&1; tail -2 gpurun_out/ncu_k1_r2.log
echo "== wave-cap sweep (lag 2, growth 2)"
for w in 8192 16384 32768; do
  timeout 300 python bench.py --wave $w --steps 8 --warmup 2 --no-cpu --plan-leg 0 > gpurun_out/bench_r2_w$w.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_w$w.json')); print('wave $w', round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,2), 'waves', d['config']['waves_per_step'], 'yield', round(d['config']['yield'],3), 'depth', d['config']['plan_depth_edges'], {k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()})" || true
done
echo "== growth 1 (Planner default) for reference"
timeout 300 python bench.py --growth 1 --steps 8 --warmup 2 --no-cpu --plan-leg 0 > gpurun_out/bench_r2_g1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_g1.json')); print('growth 1', round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,2), 'waves', d['config']['waves_per_step'])" || true
echo "== C5 full: FD-Riccati provider, 1M nodes, one plan"
timeout 600 python bench.py --config c5 --steps 1 --warmup 0 --cpu-seconds 5 --plan-leg 0 > gpurun_out/bench_r2_c5.json 2> gpurun_out/bench_r2_c5.err || tail -c 600 gpurun_out/bench_r2_c5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_c5.json')); print('c5', round(d['value']/1e6,4),'M nodes/s', round(d['ms_per_step'],1),'ms/plan', 'waves', d['config']['waves_per_step'], 'yield', round(d['config']['yield'],4), d['stage_ms'])" || true
echo "== ncu --set full: Riccati rollout + metric kernels (C5 at 20k nodes)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'steer_x|metric_coop' -s 30 -c 4 -o gpurun_out/ncu_riccati_r2 -f python tools/c5_once.py > gpurun_out/ncu_riccati_r2.log 2>&1; tail -2 gpurun_out/ncu_riccati_r2.log
date
