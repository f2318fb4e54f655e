"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Recipe for oracle/_ref/: a verbatim, unmodified copy of the reference's library and demo files, taken from
where they lie under /root/reference, so that the REAL reference (not the port in oracle/seqplanner.py) can be
timed on the GPU box's host cores by bench.py (`cpu_baseline.kind = "reference"`, `--impl reference`) and its
real demo model sections can be fed to the recognition tests.  oracle/_ref/ is git-ignored (it never enters the
history) but not gpurun-ignored (it travels with a snapshot, like the built .so).  Run by __graft_entry__.build()
whenever /root/reference is present; a box without oracle/_ref falls back to the port and says so.

The reference is Python 2 with implicit relative imports (lqrrt/__init__.py:1-2); `load()` below imports the
copy unmodified on Python 3 the way tests/ref_shim.py imports the original.
"""
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference'
REF_DST = os.path.join(HERE, '_ref')
DEMOS = ('demo_boat_novice', 'demo_boat_intermediate', 'demo_boat_advanced', 'demo_car')


def build():
    """Copy the files (no edits).  Returns True if oracle/_ref is complete afterwards."""
    if os.path.isfile(os.path.join(REF_SRC, 'lqrrt', 'planner.py')):
        os.makedirs(os.path.join(REF_DST, 'lqrrt'), exist_ok=True)
        os.makedirs(os.path.join(REF_DST, 'demos'), exist_ok=True)
        for f in ('__init__.py', 'planner.py', 'tree.py', 'constraints.py'):
            shutil.copyfile(os.path.join(REF_SRC, 'lqrrt', f), os.path.join(REF_DST, 'lqrrt', f))
        for d in DEMOS:
            shutil.copyfile(os.path.join(REF_SRC, 'demos', d + '.py'), os.path.join(REF_DST, 'demos', d + '.py'))
    return available()


def available():
    return os.path.isfile(os.path.join(REF_DST, 'lqrrt', 'planner.py')) and \
        all(os.path.isfile(os.path.join(REF_DST, 'demos', d + '.py')) for d in DEMOS)


def load():
    """import the copied reference package as module `reference_lqrrt_copy` (never shadows this repo's `lqrrt`)."""
    if 'reference_lqrrt_copy' in sys.modules:
        return sys.modules['reference_lqrrt_copy']
    pkg_dir = os.path.join(REF_DST, 'lqrrt')
    saved_path = list(sys.path)
    saved = {k: sys.modules.pop(k) for k in ('lqrrt', 'tree', 'constraints', 'planner') if k in sys.modules}
    try:
        sys.path.insert(0, pkg_dir)
        spec = importlib.util.spec_from_file_location('reference_lqrrt_copy', os.path.join(pkg_dir, '__init__.py'),
                                                      submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules['reference_lqrrt_copy'] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k in ('tree', 'constraints', 'planner'):
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules['reference_copy_' + k] = m
        sys.modules.update(saved)
    return mod


def demo_model(name, seed=0):
    """exec a copied demo's model section (everything above its PLAN banner) after np.random.seed(seed);
    returns the namespace with the demo's own dynamics / lqr / erf / is_feasible and parameters."""
    import numpy as np
    ref = load()
    path = os.path.join(REF_DST, 'demos', name + '.py')
    src = open(path).read()
    src = src[:src.index('################################################# PLAN')]
    sys.modules['lqrrt_demo_shadow_copy'] = ref
    src = src.replace('import lqrrt', 'import lqrrt_demo_shadow_copy as lqrrt')
    ns = {'__name__': 'demo_model_' + name}
    np.random.seed(seed)
    exec(compile(src, path, 'exec'), ns)
    return ns


if __name__ == '__main__':
    print('oracle/_ref complete:', build())
