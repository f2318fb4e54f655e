import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
DEMOS = ['demo_boat_novice', 'demo_car', 'demo_boat_intermediate', 'demo_boat_advanced']


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def problem(name, seed=0):
    from lqrrt_b200 import configs
    return configs.BY_NAME[name](seed)


def setup_draws(name):
    """Uniforms the demo's obstacle-grid setup consumes from the global stream before planning
    (36 slots x 2, demo_boat_novice.py:153); demo_car's default obstacle choice draws none."""
    return 0 if name == 'demo_car' else 72


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return 0
