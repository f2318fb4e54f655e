#!/bin/bash
# Round 2, call 19 (1 GPU): final evidence run -- full GPU suite, smoke, bench line for every config, reference arm, ncu
# launch list + ncu --set full captures (csv exports), compute-sanitizer.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call19.log) 2>&1
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
print(round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2), d['e2e'].get('host_ms_per_step'), 'digest', d['tree_digest'], 'clocks', d['clocks'])
print('plan', d.get('plan')); print('cpu', d.get('cpu_baseline')); r=d['roofline']; print('roofline frac', r['frac'], 'fma', r['fma_pipe_frac'], 'alone', r.get('alone'), 'in pipeline', r['in_pipeline']['frac'], 'stage', d['stage_ms'])
PY
echo "== other configs"
for c in c3 c2 c1 c5k c5; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 1 --cpu-seconds 5 --plan-leg 0 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err || tail -c 600 gpurun_out/bench_r2_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$c.json')); print('$c', round(d['value']/1e6,3),'M nodes/s', round(d['ms_per_step'],2),'ms/plan e2e', round(d['e2e']['value']/1e6,3), 'waves', d['config']['waves_per_step'], 'yield', round(d['config']['yield'],3), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], {k:round(v/d['steps'],1) for k,v in d['stage_ms'].items()})" || true
done
echo "== reference arm"
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 8 | tail -1 > gpurun_out/bench_r2_ref.json; cut -c1-300 gpurun_out/bench_r2_ref.json
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 175 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu --plan-leg 0 > gpurun_out/launches_r2.log 2>&1; tail -1 gpurun_out/launches_r2.log | cut -c1-160
cap() { name=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -f -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1; tail -1 gpurun_out/$name.log | cut -c1-160
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ls -la gpurun_out/${name}_raw.csv
}
echo "== ncu --set full: scan (packed, scalar) at n = 6, W = 8192, N = 1e5"
cap ncu_k1_r2 -k regex:nearest_diag -c 2 python tools/k1_once.py
echo "== ncu --set full: rollout + commit kernels of a C4 plan (saturated waves)"
cap ncu_chain_r2 -k regex:'steer_x|commit_x' -s 50 -c 4 python bench.py --steps 1 --warmup 0 --no-cpu --plan-leg 0
echo "== ncu --set full: C5 (whitened search, gain memo), 20k nodes, waves of 2048"
cap ncu_riccati_r2b -k regex:'steer_x|whiten|nearest_diag' -s 24 -c 6 python tools/c5_once.py
echo "== compute-sanitizer"
PIPE_NODES=4000 PIPE_WAVE=256 LQRRT_LAG=2 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_r2.log python tools/pipe_once.py | tail -1; tail -2 gpurun_out/sanitizer_memcheck_r2.log
PIPE_NODES=1500 PIPE_WAVE=128 LQRRT_LAG=2 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_r2.log python tools/pipe_once.py | tail -1; tail -2 gpurun_out/sanitizer_racecheck_r2.log
C5_NODES=3000 C5_WAVE=256 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_c5_r2.log python tools/c5_once.py | tail -1; tail -2 gpurun_out/sanitizer_memcheck_c5_r2.log
C5_NODES=600 C5_WAVE=64 timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_c5_r2.log python tools/c5_once.py | tail -1; tail -2 gpurun_out/sanitizer_racecheck_c5_r2.log
du -sh gpurun_out
date
