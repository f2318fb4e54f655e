#!/bin/bash
# Round 2, call 8 (2 GPUs): pipelined waves (lag 2), folded partial minima; tests + 1/2-GPU bench for both lags.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call8.log) 2>&1
date
echo "== pipeline / wave / planner tests"
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_wave.py tests/test_gpu_planner.py tests/test_gpu_nearest.py -q --timeout 600 2>&1 | tail -15
echo "== multi-GPU tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 2>&1 | tail -6
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
run() { n=$1; shift
  if [ "$n" = 1 ]; then env "$@" timeout 300 python bench.py --gpus 1 --steps 8 --warmup 2 --no-cpu --plan-leg 0 2>&1 | tail -1 > gpurun_out/t.json
  else env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 8 --warmup 2 --no-cpu --plan-leg 0 2>&1 | tail -1 > gpurun_out/t.json; fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/t.json'))
    print('N=%d $*'%d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2),' stages/plan(ms):',{k:round(v/d['steps'],2) for k,v in d['stage_ms'].items()},'host',d['e2e'].get('wave_loop_host_ms_per_step'),'waves',d['config']['waves_per_step'],'digest',d.get('tree_digest'))
except Exception as e:
    print('N=$n unreadable', e, open('gpurun_out/t.json').read()[-1500:])
PY
}
run 1 LQRRT_LAG=1
run 1 LQRRT_LAG=2
run 2 LQRRT_LAG=1
run 2 LQRRT_LAG=2
echo "== chain kernels at W = 1024, lag 1 and 2"
LQRRT_LAG=1 PIPE_PLANS=3 timeout 120 python tools/pipe_once.py
LQRRT_LAG=2 PIPE_PLANS=3 timeout 120 python tools/pipe_once.py
echo "== goal-reaching policy (W <= 512), lag 1 and 2"
for lag in 1 2; do LQRRT_LAG=$lag timeout 200 python - <<'PY'
import os, time, numpy as np, sys
sys.path.insert(0, '.')
from lqrrt_b200 import configs
prob = configs.c4_boat_advanced(0)
pl = prob.make_planner(dtype='f32', wave_max=512, wave_growth=1.0, max_nodes=prob.max_nodes, min_time=0.0, max_time=10.0)
ms=[]; T=[]; w=[]
for s in range(9):
    pl.seed = 5000 + s
    t0=time.perf_counter(); pl.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias); dt=time.perf_counter()-t0
    if s and pl.plan_reached_goal: ms.append(dt*1e3); T.append(pl.T); w.append(pl.stats['waves'])
print('lag', os.environ['LQRRT_LAG'], 'reached %d/8'%len(ms), 'ms to first plan', np.round(np.mean(ms),1) if ms else None, 'T', np.round(np.mean(T),1) if T else None, 'waves', np.mean(w) if w else None)
PY
done
date
