#!/bin/bash
# Round 2, call 13 (1 GPU): evidence run on the final build -- full GPU suite, bench line for every config, sampler-ahead
# A/B at the multi-GPU working point, ncu launch list + ncu --set full captures (converted to csv on the box: the
# reports themselves are too large to bring back), compute-sanitizer.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call13.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench c4 (default command)"
timeout 600 python bench.py > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; tail -c 400 gpurun_out/bench_r2_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_c4.json'))
print(round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2), d['e2e'].get('host_ms_per_step'), 'digest', d['tree_digest'])
print('plan', d.get('plan')); print('cpu', d.get('cpu_baseline')); print('roofline frac', d['roofline']['frac'], 'fma', d['roofline']['fma_pipe_frac'], 'stage', d['stage_ms'])
PY
one() { tag=$1; shift
  timeout 600 python bench.py --no-cpu --plan-leg 0 "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err || tail -c 400 gpurun_out/bench_r2_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$tag.json')); print('$tag', round(d['value']/1e6,3),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,3), d['e2e']['host_ms_per_step'], 'waves', d['config']['waves_per_step'], {k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()})" || true
}
echo "== c4 variants"
one c4_w8192 --wave 8192
one c4_lag1 --lag 1
LQRRT_SAMPLE_AHEAD=0 one c4_noahead
echo "== other configs"
for c in c3 c2 c1 c5k; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 1 --cpu-seconds 5 --plan-leg 0 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err || tail -c 600 gpurun_out/bench_r2_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$c.json')); print('$c', round(d['value']/1e6,3),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,3), 'waves', d['config']['waves_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])" || true
done
timeout 600 python bench.py --config c5 --steps 1 --warmup 0 --cpu-seconds 5 --plan-leg 0 > gpurun_out/bench_r2_c5.json 2> gpurun_out/bench_r2_c5.err || tail -c 600 gpurun_out/bench_r2_c5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_c5.json')); print('c5', round(d['value']/1e6,4),'M nodes/s', round(d['ms_per_step'],1),'ms/plan', 'waves', d['config']['waves_per_step'], 'yield', round(d['config']['yield'],4), d['stage_ms'], 'cpu', d['cpu_baseline'])" || true
echo "== reference arm"
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 8 | tail -1 > gpurun_out/bench_r2_ref.json; cut -c1-300 gpurun_out/bench_r2_ref.json
echo "== sampler ahead of the chain, fixed waves of 1024 / 2048 (one GPU's share of an 8-GPU wave), lag 2"
for w in 1024 2048; do for a in 0 1; do
  echo "wave $w ahead $a: $(PIPE_NODES=100000 PIPE_WAVE=$w PIPE_PLANS=3 LQRRT_LAG=2 LQRRT_SAMPLE_AHEAD=$a timeout 120 python tools/pipe_once.py | tail -1)"
done; done
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 170 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu --plan-leg 0 > gpurun_out/launches_r2.log 2>&1; tail -1 gpurun_out/launches_r2.log | cut -c1-160
cap() { name=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -f -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1; tail -1 gpurun_out/$name.log | cut -c1-160
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/${name}_details.csv 2>/dev/null
  ls -la gpurun_out/${name}_raw.csv
}
echo "== ncu --set full: scan (packed, scalar) at n = 6, W = 8192, N = 1e5"
cap ncu_k1_r2 -k regex:nearest_diag -c 2 python tools/k1_once.py
echo "== ncu --set full: rollout + commit kernels of a C4 plan (saturated waves)"
cap ncu_chain_r2 -k regex:'steer_x|commit_x' -s 50 -c 4 python bench.py --steps 1 --warmup 0 --no-cpu --plan-leg 0
echo "== ncu --set full: Riccati rollout + metric kernels (C5, 20k nodes)"
cap ncu_riccati_r2 -k regex:'steer_x|metric_coop|nearest_dense' -s 24 -c 6 python tools/c5_once.py
echo "== compute-sanitizer (lag 2, sampler ahead; then the indexed search on)"
PIPE_NODES=4000 PIPE_WAVE=256 LQRRT_LAG=2 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_r2.log python tools/pipe_once.py | tail -1; tail -2 gpurun_out/sanitizer_memcheck_r2.log
PIPE_NODES=1500 PIPE_WAVE=128 LQRRT_LAG=2 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_r2.log python tools/pipe_once.py | tail -1; tail -2 gpurun_out/sanitizer_racecheck_r2.log
PIPE_NODES=4000 PIPE_WAVE=256 LQRRT_LAG=2 LQRRT_GRID=1 LQRRT_GRID_MIN=0 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_grid_r2.log python tools/pipe_once.py | tail -1; tail -2 gpurun_out/sanitizer_memcheck_grid_r2.log
du -sh gpurun_out
date
