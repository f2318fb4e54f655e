"""Host mirror of the reference interface: constructor / setter validation and messages
(planner.py:468-592, constraints.py:39-61) -- no GPU needed until update_plan is called."""
import numpy as np
import pytest

import lqrrt_b200 as lqrrt
from lqrrt_b200 import models, configs


def _planner(**kw):
    prob = configs.demo_boat_novice(0)
    return prob, prob.make_planner(**kw)


def test_package_surface():
    import lqrrt as alias
    assert alias.Planner is lqrrt.Planner and alias.Constraints is lqrrt.Constraints
    assert hasattr(lqrrt, 'Tree')


def test_constraints_validation():
    feas = models.Feasibility()
    c = lqrrt.Constraints(2, 1, [-1, np.inf], feas)
    assert c.goal_buffer.dtype == np.float64 and c.goal_buffer[0] == 1.0
    with pytest.raises(ValueError, match="goal_buffer must have same dimensionality"):
        c.set_buffers([1, 2, 3])
    with pytest.raises(ValueError, match="Expected is_feasible to be a function"):
        c.set_feasibility_function(3)


def test_planner_setters_match_reference_rules():
    prob, p = _planner()
    assert p.horizon_iters == 20 and p.hfactor == 0 and p.nstates == 6 and p.ncontrols == 3
    with pytest.raises(ValueError, match="min_time must be less than or equal"):
        p.set_runtime(min_time=5, max_time=1)
    p.set_runtime(min_time=1, max_time=5)
    with pytest.raises(ValueError, match="Expected sys_time to be a function"):
        p.set_runtime(sys_time=3)
    with pytest.raises(ValueError, match="horizon must be at least as big as dt"):
        p.set_resolution(horizon=0.01)
    p.set_resolution(horizon=2)
    with pytest.raises(ValueError, match="Shape of error_tol"):
        p.set_resolution(error_tol=[1, 2])
    p.set_resolution(horizon=(0.1, 3))
    assert p.hfactor == 2 and p.horizon_iters == 1 and list(p.hspan) == [1, 30]
    with pytest.raises(ValueError, match="min < max"):
        p.set_resolution(horizon=(3, 1))
    p.set_resolution(horizon=2)
    with pytest.raises(ValueError, match="goal state must have same dimensionality"):
        p.set_goal([1, 2])
    p.set_goal(None)
    assert p.goal is None
    p.dt = 0.05                      # plain attribute, as lqrrt_node.py:259 assigns it
    assert p.dt == 0.05
    p.kill_update(); assert p.killed
    p.unkill(); assert not p.killed


def test_plain_callables_are_rejected_loudly():
    """No CPU fallback: arbitrary Python closures cannot run in a kernel."""
    prob = configs.demo_boat_novice(0)
    cons = lqrrt.Constraints(6, 3, prob.goal_buffer, lambda x, u: True)
    with pytest.raises(TypeError, match="no CPU fallback"):
        lqrrt.Planner(lambda x, u, dt: x, lambda x, u: (np.eye(6), np.zeros((3, 6))), cons, horizon=2, dt=0.1)
    with pytest.raises(ValueError, match="Expected dynamics to be a function"):
        lqrrt.Planner(None, prob.system.lqr, cons, horizon=2, dt=0.1)


def test_model_blobs_have_the_wire_sizes():
    for name, n_sys in (('demo_boat_novice', 46), ('demo_car', 9), ('c1', 5), ('c5', 2)):
        s = configs.BY_NAME[name](0).system
        assert len(s.dynamics.blob()) == n_sys
        assert s.erf.mask(s.nstates) in (0, 1, 4)
    car = configs.demo_car(0).system
    assert car.is_feasible.prims[0].flags == 1 and car.is_feasible.prims[0].verts.shape[0] == 2
    nov = configs.demo_boat_novice(0).system
    assert len(nov.is_feasible.prims[0].obs) < 36        # sentinel slots dropped


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from lqrrt_b200 import _native
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        _native.load_library()


def test_engine_option_arguments(monkeypatch):
    """Issue-level options (lq_set_option): validated on the host, results never depend on them."""
    from lqrrt_b200 import planner as pl
    monkeypatch.delenv('LQRRT_NEAREST', raising=False)
    monkeypatch.delenv('LQRRT_PREFETCH', raising=False)
    prob, p = _planner()
    assert p.nearest_kernel == pl.DEFAULT_NEAREST_KERNEL and p.prefetch == pl.DEFAULT_PREFETCH
    assert p.prefetch_active == pl.DEFAULT_PREFETCH
    with pytest.raises(ValueError, match="nearest_kernel must be one of"):
        _planner(nearest_kernel='fastest')
    _, q = _planner(nearest_kernel='scalar', prefetch=False)
    assert q.nearest_kernel == 'scalar' and not q.prefetch_active
    _, r = _planner(prefetch=True, pipeline=False)          # nothing to overlap with without the pipelined loop
    assert not r.prefetch_active
    monkeypatch.setenv('LQRRT_NEAREST', 'two_level')
    monkeypatch.setenv('LQRRT_PREFETCH', '0')
    _, e = _planner()
    assert e.nearest_kernel == 'two_level' and not e.prefetch
    # multi-GPU planners take the prefetch too (host-scheduled waves; the device-scheduled loop runs every wave's
    # sampler on its scan stream anyway)
    monkeypatch.delenv('LQRRT_PREFETCH', raising=False)
    _, m = _planner()
    m.world = 2
    assert m.prefetch_active
    with pytest.raises(ValueError, match="schedule must be"):
        _planner(schedule='gpu')
    _, h = _planner(schedule='host', fresh_samples=False)
    assert h.schedule == 'host' and not h.fresh_samples


def test_bench_cli_defaults():
    """The driver runs `python bench.py` bare: N = 1, the C4 workload, growth-2 waves, the planner's own kernels."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(root, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        import argparse
        captured = {}
        orig = argparse.ArgumentParser.parse_args

        def grab(self, *a, **k):
            ns = orig(self, *a, **k)
            captured['ns'] = ns
            raise SystemExit(0)
        argparse.ArgumentParser.parse_args = grab
        with pytest.raises(SystemExit):
            mod.main()
    finally:
        argparse.ArgumentParser.parse_args = orig
        sys.argv = argv
    ns = captured['ns']
    assert (ns.gpus, ns.config, ns.impl, ns.wave, ns.growth, ns.nearest, ns.prefetch, ns.lag) == (1, 'c4', 'ours', 16384, 2.0, None, None, 2)
    assert ns.steps >= 1 and ns.warmup >= 3


def test_device_schedule_rule_mirrors_the_host_rule():
    """The commit kernel sizes waves with policy_wave / policy_slice (csrc/wave.cuh, exported as plain host functions):
    they must agree with lqrrt_b200.waves.wave_size / partition, which the host-scheduled loop and the multi-rank
    gloo test use, for every size the planner can see -- no GPU needed."""
    import ctypes as C
    from lqrrt_b200 import _native, waves
    lib = _native.load_library()
    rng = np.random.RandomState(0)
    for _ in range(4000):
        n = int(rng.choice([1, 2, 33, 100, 4095, 4096, 4097, 10 ** 5, 10 ** 6, int(rng.randint(1, 2 * 10 ** 6))]))
        growth = float(rng.choice([0.5, 1.0, 2.0, 3.7, 4.0]))
        wmax = int(rng.choice([32, 512, 8192, 65536]))
        fixed = int(rng.choice([0, 0, 0, 1, 64, 1000]))
        want = waves.wave_size(n, fixed or None, wave_min=32, wave_max=wmax, growth=growth)
        assert lib.lq_policy_wave(fixed, 32, wmax, growth, n) == want
        world = int(rng.choice([1, 2, 3, 4, 8])); rank = int(rng.randint(world))
        f = C.c_int32(); c = C.c_int32()
        assert lib.lq_policy_slice(want, rank, world, C.byref(f), C.byref(c)) == 0
        assert (f.value, c.value) == waves.partition(want, rank, world)
    # slices tile the wave in rank order
    for W in (0, 1, 7, 8192, 8191):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                first, count = waves.partition(W, r, world)
                assert first == pos
                pos += count
            assert pos == W


def test_reference_arm_prints_the_gpu_arms_config():
    """Both arms of bench.py carry the same `config` object (what is measured per run lives under `run`), and the reference
    arm is a bounded sample: one line, metric / unit / direction of the GPU arm, no GPU touched."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--cpu-seconds', '2'], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'tree_expansions_per_s' and d['unit'] == 'nodes/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['gpu_launches'] == 0
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod2', os.path.join(root, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import argparse
    ns = argparse.Namespace(wave=16384, scaling='strong', lag=2, nearest=None, growth=2.0, schedule=None, config='c4')
    assert d['config'] == mod.static_config(ns, mod.make_problem('c4'), 1)
