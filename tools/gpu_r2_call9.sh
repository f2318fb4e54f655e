#!/bin/bash
# Round 2, call 9 (1 GPU): evidence run -- tests touched since the last full suite, full bench line for every config,
# ncu launch list of the bench command, ncu --set full of the scan / rollout / commit kernels, compute-sanitizer.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call9.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_planner.py tests/test_gpu_recognize.py -q --timeout 600 2>&1 | tail -6
echo "== bench c4 (default command)"
timeout 600 python bench.py > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; tail -c 400 gpurun_out/bench_r2_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_c4.json'))
print(round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2), d['e2e'].get('host_ms_per_step'), 'digest', d['tree_digest'])
print('plan', d.get('plan')); print('cpu', d.get('cpu_baseline')); print('roofline frac', d['roofline']['frac'], 'fma', d['roofline']['fma_pipe_frac'])
PY
echo "== bench c4 lag 1"
timeout 300 python bench.py --lag 1 --no-cpu --plan-leg 0 > gpurun_out/bench_r2_c4_lag1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_c4_lag1.json')); print(round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,2))"
for c in c3 c2 c1 c5k; do
  echo "== bench $c"
  timeout 600 python bench.py --config $c --steps 4 --warmup 1 --cpu-seconds 5 --plan-leg 0 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err || tail -c 600 gpurun_out/bench_r2_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$c.json')); print('$c', round(d['value']/1e6,3),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,3), 'waves', d['config']['waves_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])" || true
done
echo "== reference arm"
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 8 | tail -1 > gpurun_out/bench_r2_ref.json; cut -c1-300 gpurun_out/bench_r2_ref.json
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 2 --no-cpu --plan-leg 0 > gpurun_out/launches_r2.log 2>&1; tail -1 gpurun_out/launches_r2.log | cut -c1-200
echo "== ncu --set full: scan (saturated launches), rollout and commit kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nearest_diag_pk' -s 70 -c 2 -o gpurun_out/ncu_nearest_r2 python bench.py --steps 1 --warmup 0 --no-cpu --plan-leg 0 > gpurun_out/ncu_nearest_r2.log 2>&1; tail -1 gpurun_out/ncu_nearest_r2.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'steer_x|commit_x' -s 70 -c 4 -o gpurun_out/ncu_chain_r2 python bench.py --steps 1 --warmup 0 --no-cpu --plan-leg 0 > gpurun_out/ncu_chain_r2.log 2>&1; tail -1 gpurun_out/ncu_chain_r2.log | cut -c1-200
echo "== compute-sanitizer"
PIPE_NODES=4000 PIPE_WAVE=256 LQRRT_LAG=2 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_r2.log python tools/pipe_once.py | tail -1; tail -3 gpurun_out/sanitizer_memcheck_r2.log
PIPE_NODES=1500 PIPE_WAVE=128 LQRRT_LAG=2 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_r2.log python tools/pipe_once.py | tail -1; tail -3 gpurun_out/sanitizer_racecheck_r2.log
date
