"""
Loads the UNMODIFIED reference from /root/reference (only present in the build
container, never on the GPU box) for pinning the oracle and for generating
tests/golden/ (SURVEY.md section 8c recipe).  Nothing under `-m gpu`, smoke() or bench.py
imports this file.
"""
import os
import sys

import numpy as np

REF_ROOT = '/root/reference'


def have_reference():
    return os.path.isfile(os.path.join(REF_ROOT, 'lqrrt', 'planner.py'))


def import_reference():
    """The reference is Python 2 with implicit relative imports
    (lqrrt/__init__.py:1-2); putting both the repo root and the package dir on
    sys.path makes `import lqrrt` resolve unmodified on Python 3.  The module is
    registered as `reference_lqrrt` so it never shadows this repo's `lqrrt`."""
    if 'reference_lqrrt' in sys.modules:
        return sys.modules['reference_lqrrt']
    import importlib.util
    pkg_dir = os.path.join(REF_ROOT, 'lqrrt')
    saved_path = list(sys.path)
    saved = {k: sys.modules.pop(k) for k in ('lqrrt', 'tree', 'constraints', 'planner') if k in sys.modules}
    try:
        sys.path.insert(0, pkg_dir)
        spec = importlib.util.spec_from_file_location(
            'reference_lqrrt', os.path.join(pkg_dir, '__init__.py'), submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules['reference_lqrrt'] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k in ('tree', 'constraints', 'planner'):
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules['reference_' + k] = m
        sys.modules.update(saved)
    return mod


def load_demo_model(name, seed, overrides=None):
    """exec a demo's model section (everything above its PLAN banner) after
    np.random.seed(seed); returns the resulting namespace."""
    ref = import_reference()
    path = os.path.join(REF_ROOT, 'demos', name + '.py')
    src = open(path).read()
    cut = src.index('################################################# PLAN')
    src = src[:cut]
    if overrides:
        for a, b in overrides.items():
            assert a in src, a
            src = src.replace(a, b)
    ns = {'__name__': 'demo_model'}
    sys.modules['lqrrt_demo_shadow'] = ref
    src = src.replace('import lqrrt', 'import lqrrt_demo_shadow as lqrrt')
    np.random.seed(seed)
    exec(compile(src, path, 'exec'), ns)
    return ns


class FrameClock(object):
    """sys_time that ticks only when called from update_plan (planner.py:230,286)
    and not from _steer (:380,402): min_time = max_time = K gives K iterations."""

    def __init__(self):
        self.t = -1

    def __call__(self):
        if sys._getframe(1).f_code.co_name == 'update_plan':
            self.t += 1
        return self.t


def run_reference_demo(name, seed, iterations, planner_kwargs, overrides=None, tries=10, pruning=True):
    """SURVEY.md section 8c deterministic replay recipe."""
    ref = import_reference()
    ns = load_demo_model(name, seed, overrides)
    cons = ref.Constraints(nstates=ns['nstates'], ncontrols=ns['ncontrols'],
                           goal_buffer=ns['goal_buffer'], is_feasible=ns['is_feasible'])
    kw = dict(error_tol=ns['error_tol'], erf=ns['erf'], min_time=iterations, max_time=iterations,
              max_nodes=1E5, goal0=ns['goal'], sys_time=FrameClock(), printing=False)
    kw.update(planner_kwargs)
    planner = ref.Planner(ns['dynamics'], ns['lqr'], cons, **kw)
    planner.update_plan(ns['x0'], ns['sample_space'], goal_bias=ns['goal_bias'], xrand_gen=tries, pruning=pruning)
    return planner, ns
