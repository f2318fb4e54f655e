python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8
cat > /tmp/c5.py <<'PY'
import time, sys, numpy as np, torch
sys.path.insert(0,'.')
from lqrrt_b200 import configs
prob = configs.c5_double_integrator(0, riccati=True, max_nodes=20000)
pl = prob.make_planner(dtype='f32', seed=1, min_time=60, max_time=60)
pl.profile = True
for rep in range(2):
    t=time.time(); pl.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias); dt=time.time()-t
    print('riccati nodes', pl.tree.size, 'time %.3f s'%dt, 'nodes/s %.3g'%(pl.tree.size/dt), pl.stats['waves'], 'waves', pl.stats['samples'], 'samples', flush=True)
print(pl._engine.profile_read()['ms'])
prob = configs.c1_pendulum(0)
pl = prob.make_planner(dtype='f32', seed=1, min_time=60, max_time=60, max_nodes=100000)
for rep in range(2):
    t=time.time(); pl.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias); dt=time.time()-t
    print('c1 pendulum nodes', pl.tree.size, 'time %.3f s'%dt, 'nodes/s %.3g'%(pl.tree.size/dt), pl.stats['waves'], 'waves', 'reached', pl.plan_reached_goal, flush=True)
PY
timeout 200 python /tmp/c5.py 2>&1 | tail -6
