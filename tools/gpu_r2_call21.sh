#!/bin/bash
# Round 2, call 21 (1 GPU): what the driver runs at round end, on the final tree -- GPU suite, smoke, the bench command.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call21.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 300 gpurun_out/bench_r2_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final.json'))
print(round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2), d['e2e'].get('host_ms_per_step'), 'digest', d['tree_digest'], 'clocks', d['clocks'], 'launches', d['gpu_launches'])
print('config', d['config']); print('run', d['run']); print('plan', d.get('plan')); print('cpu', d.get('cpu_baseline')); r=d['roofline']; print('roofline', r['bound'], r['frac'], r['fma_pipe_frac'], r['alone'], r['traffic'], r['traffic_source'])
PY
date
