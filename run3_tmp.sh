python -m pytest tests/test_gpu_recognize.py -q --timeout 600 2>&1 | tail -15
LQRRT_TUNING=32 timeout 100 python - <<'PY' 2>&1 | tail -3
import time, sys, numpy as np
sys.path.insert(0,'.')
from lqrrt_b200 import configs
prob = configs.c5_double_integrator(0, riccati=True, max_nodes=20000)
pl = prob.make_planner(dtype='f32', seed=1, min_time=60, max_time=60)
pl.profile = True
for rep in range(2):
    t=time.time(); pl.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias); dt=time.time()-t
    print('riccati lanes32 nodes', pl.tree.size, 'time %.3f s'%dt, pl.stats['waves'], 'waves', flush=True)
print(pl._engine.profile_read()['ms'])
PY
