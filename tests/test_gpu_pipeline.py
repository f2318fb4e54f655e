"""
The device-scheduled wave pipeline (lq_plan_run: waves sized on the device, old-tree scan of wave k+1 under wave
k's rollouts, per-query check of the ignore marks, results exchanged uncompacted) must grow exactly the tree of
the host-scheduled stepwise calls, which tests/test_gpu_wave.py pins node for node to oracle/wave.py.
"""
import numpy as np
import pytest

from conftest import problem
from lqrrt_b200 import _native as nat

pytestmark = pytest.mark.gpu


def _grow(prob, dtype, schedule, opts=None, goal_bias=None, **kw):
    args = dict(dtype=dtype, seed=11, max_nodes=3000, min_time=1e9, max_time=1e9, wave_max=512, schedule=schedule, lag=1)
    args.update(kw)
    planner = prob.make_planner(**args)
    planner.engine_options = dict(opts or {})
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias if goal_bias is None else goal_bias)
    t = planner.tree
    out = dict(ok=ok, size=t.size, state=np.array(t.state), pID=list(t.pID), cost=np.array(t.cost),
               waves=planner.stats['waves'], samples=planner.stats['samples'], reached=planner.plan_reached_goal,
               T=planner.T, schedule=planner.stats['schedule'])
    if planner.plan_reached_goal:
        out['node_seq'] = list(planner.node_seq)
        out['x_seq'] = np.array(planner.x_seq)
    planner._engine.close()
    return out


def _same(a, b):
    assert a['size'] == b['size'] and a['pID'] == b['pID']
    assert a['state'].tobytes() == b['state'].tobytes()
    assert np.array_equal(a['cost'], b['cost'])
    # (the host-scheduled loop runs one wave ahead of the counters it reads, so it may have launched one more
    # wave onto the already full tree; that wave appends nothing)
    assert abs(a['waves'] - b['waves']) <= 1
    assert a['reached'] == b['reached'] and a['T'] == b['T']
    if a['reached']:
        assert a['node_seq'] == b['node_seq'] and a['x_seq'].tobytes() == b['x_seq'].tobytes()


@pytest.mark.parametrize('name,dtype', [('demo_boat_novice', 'f64'), ('demo_boat_novice', 'f32'), ('demo_car', 'f32'),
                                        ('c4', 'f32'), ('c1', 'f32'), ('c1', 'f64')])
def test_device_schedule_grows_the_host_schedule_tree(cuda_device, name, dtype):
    prob = problem(name)
    host = _grow(prob, dtype, 'host')
    dev = _grow(prob, dtype, 'device')
    assert host['schedule'] == 'host' and dev['schedule'] == 'device'
    assert dev['size'] == 3001 and dev['ok'] is False           # planner.py:330-336: over max_nodes returns False
    _same(host, dev)
    # ... with the speculative scan off (one whole-tree scan per wave on the chain stream)
    _same(host, _grow(prob, dtype, 'device', {nat.LQ_OPT_SPECULATIVE_SCAN: 0}))
    # ... with every query forced through the rollout kernel's own rescan (the path a stale winner takes)
    _same(host, _grow(prob, dtype, 'device', {nat.LQ_OPT_FORCE_RESCAN: 1}, max_nodes=3000))
    # ... with the obstacle tables read from global memory instead of the shared-memory copies
    _same(host, _grow(prob, dtype, 'device', {nat.LQ_OPT_STAGE_TABLES: 0}))
    # ... and with both rollout CTA sizes
    _same(host, _grow(prob, dtype, 'device', {nat.LQ_OPT_STEER_BLOCK: 32}))
    _same(host, _grow(prob, dtype, 'device', {nat.LQ_OPT_STEER_BLOCK: 128}))


@pytest.mark.parametrize('pruning_bias', [0.5, 0.9])
def test_goal_reaching_plans_with_ignore_marks_agree(cuda_device, pruning_bias):
    """Once goal nodes exist every commit adds ignore marks while the next wave's old-tree scan is already
    running: the per-query check + rescan must still select what the stepwise scan (which sees the final
    marks) selects."""
    prob = problem('demo_boat_novice')
    gb = [pruning_bias, pruning_bias, 0, 0, 0, 0]
    host = _grow(prob, 'f32', 'host', goal_bias=gb, max_nodes=6000, wave_max=256)
    dev = _grow(prob, 'f32', 'device', goal_bias=gb, max_nodes=6000, wave_max=256)
    assert host['reached'] and dev['reached']
    _same(host, dev)


@pytest.mark.parametrize('lag', [1, 2])
@pytest.mark.parametrize('name,W,nodes,bias', [('demo_boat_novice', 48, 400, None), ('demo_boat_novice', 64, 4000, [0.9, 0.9, 0, 0, 0, 0]),
                                               ('c1', 40, 600, None)])
def test_fixed_waves_match_the_wave_oracle(cuda_device, lag, name, W, nodes, bias):
    """Node-for-node against oracle/wave.py through the device-scheduled driver (float64, fixed waves), for both wave
    semantics: lag 1 (wave k searches the tree of wave k-1) and lag 2 (pipelined waves: the tree AND the ignore marks of
    wave k-2).  The strongly goal-biased case reaches the goal early, so most of its waves run with ignore marks being
    added by every commit -- the case the two copies of the marks exist for."""
    from oracle.wave import WaveOracle
    prob = problem(name)
    gb = prob.goal_bias if bias is None else bias
    planner = prob.make_planner(dtype='f64', wave=W, seed=5, max_nodes=nodes, min_time=1e9, max_time=1e9, schedule='device', lag=lag)
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=gb, xrand_gen=10)
    assert planner.stats['schedule'] == 'device' and planner.stats['lag'] == lag
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol,
                     node_capacity=int(planner.max_nodes) + 1)
    orc.reset(prob.x0)
    for w in range(planner.stats['waves']):
        X, _ = orc.sample_wave(5, w, 0, W, prob.sample_space, gb, 10)
        orc.expand_wave(X, lag=lag)
    assert planner.tree.size == orc.size
    assert planner.tree.pID == orc.parent
    tol = 1e-6 if prob.system.lqr.kind == 3 else 1e-9
    assert np.abs(planner.tree.state - orc.state).max() < tol
    if bias is not None:
        assert len(orc.goal_nodes) > 3 and sum(orc.ignore) > 10           # the marks were exercised
        assert planner.plan_reached_goal and planner.T == orc.T


@pytest.mark.parametrize('name,dtype', [('demo_boat_novice', 'f32'), ('c4', 'f32'), ('demo_car', 'f64')])
def test_pipelined_waves_do_not_depend_on_issue_options(cuda_device, name, dtype):
    """lag 2: the same tree with every query forced through the rollout kernel's rescan, with the tables read from
    global memory, and with both rollout CTA sizes (adaptive waves, pruning on)."""
    prob = problem(name)
    base = _grow(prob, dtype, 'device', lag=2)
    assert base['size'] == 3001
    _same(base, _grow(prob, dtype, 'device', {nat.LQ_OPT_FORCE_RESCAN: 1}, lag=2))
    _same(base, _grow(prob, dtype, 'device', {nat.LQ_OPT_STAGE_TABLES: 0}, lag=2))
    _same(base, _grow(prob, dtype, 'device', {nat.LQ_OPT_STEER_BLOCK: 32}, lag=2))
    _same(base, _grow(prob, dtype, 'device', {nat.LQ_OPT_STEER_BLOCK: 128}, lag=2))
    # ... with the sampler on the wave's own chain instead of ahead of it on the side stream, and other scan grid targets
    _same(base, _grow(prob, dtype, 'device', {nat.LQ_OPT_SAMPLE_AHEAD: 0}, lag=2))
    _same(base, _grow(prob, dtype, 'device', {nat.LQ_OPT_NEAREST_CTAS_PER_SM: 3}, lag=2))
    # ... and with larger waves, where the policy reaches its top size and the sampler really runs ahead
    big = _grow(prob, dtype, 'device', lag=2, max_nodes=12000, wave_max=2048)
    _same(big, _grow(prob, dtype, 'device', {nat.LQ_OPT_SAMPLE_AHEAD: 0}, lag=2, max_nodes=12000, wave_max=2048))
    gb = [0.9, 0.9] + [0] * (prob.nstates - 2)
    a = _grow(prob, dtype, 'device', goal_bias=gb, lag=2, max_nodes=5000, wave_max=128)
    b = _grow(prob, dtype, 'device', {nat.LQ_OPT_FORCE_RESCAN: 1}, goal_bias=gb, lag=2, max_nodes=5000, wave_max=128)
    _same(a, b)


def test_replans_draw_fresh_samples_unless_asked_not_to(cuda_device):
    prob = problem('demo_boat_novice')
    trees = []
    for fresh in (True, False):
        planner = prob.make_planner(dtype='f32', seed=3, max_nodes=500, min_time=1e9, max_time=1e9, wave_max=128,
                                    fresh_samples=fresh)
        pair = []
        for _ in range(2):
            planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
            pair.append(np.array(planner.tree.state).tobytes())
        trees.append(pair)
        planner.seed = 3                      # assigning the key restarts its stream
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
        assert np.array(planner.tree.state).tobytes() == pair[0]
    assert trees[0][0] != trees[0][1]          # the reference's global stream moves on between plans (planner.py:204)
    assert trees[1][0] == trees[1][1]
    assert trees[0][0] == trees[1][0]


def test_time_based_stop_and_kill(cuda_device):
    import threading
    import time
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f32', wave_max=256, min_time=0.2, max_time=0.2, max_nodes=1e6, schedule='device')
    t0 = time.time()
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=[0] * 6)
    assert ok and time.time() - t0 < 5 and planner.stats['schedule'] == 'device'
    assert planner.tree.size > 100 and len(planner.x_seq) >= 1            # timed out: closest-to-guide plan
    threading.Timer(0.2, planner.kill_update).start()
    planner.set_runtime(min_time=30, max_time=30)
    ok = planner.update_plan(prob.x0, prob.sample_space, goal_bias=[0] * 6)
    assert ok is False and planner.killed is False and planner.stats['elapsed'] < 10


def test_tree_nearest_matches_oracle(cuda_device):
    """Tree.nearest = _costs_to_go + selection (planner.py:239-247, 340-350) through lq_nearest."""
    from oracle.wave import WaveOracle
    prob = problem('demo_boat_novice')
    planner = prob.make_planner(dtype='f64', seed=2, max_nodes=1500, min_time=1e9, max_time=1e9, wave_max=256)
    planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
    tree = planner.tree
    orc = WaveOracle(prob.system, prob.goal, prob.goal_buffer, prob.horizon, prob.dt, prob.FPR, prob.error_tol)
    orc.load(tree.state, tree.pID)
    rng = np.random.RandomState(0)
    ss = np.array(prob.sample_space)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * rng.rand(40, 6)
    ids, costs = tree.nearest(Q)
    oid, oc = orc.nearest(Q, use_ignore=False)
    assert np.array_equal(ids, oid) and np.allclose(costs, oc, rtol=1e-12, atol=1e-12)
    one = tree.nearest(Q[3])
    assert one[0] == oid[3] and abs(one[1] - oc[3]) < 1e-12
