"""The C-ABI library loads on a CPU-only box and exports every function include/lqrrt_b200.h declares
(no compute calls here), and the ctypes binding declares the same set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'lqrrt_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(lq_[a-z_0-9]+)\s*\(', src)))


def test_header_declares_the_documented_surface():
    names = declared_functions()
    for must in ('lq_create', 'lq_destroy', 'lq_nearest', 'lq_steer', 'lq_wave_expand', 'lq_wave_commit',
                 'lq_wave_records', 'lq_closest', 'lq_download_nodes', 'lq_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from lqrrt_b200 import _native
    if not os.path.isfile(_native.LIB_PATH):
        pytest.fail("liblqrrt_b200.so is not built; run __graft_entry__.build()")
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.lq_abi_version() == _native.LQ_ABI_VERSION


def test_binding_covers_header():
    from lqrrt_b200 import _native
    assert sorted(_native.SYMBOLS) == declared_functions()


def test_create_fails_loudly_without_gpu_or_with_bad_config():
    """No CPU fallback: on a box without a CUDA device lq_create must return an error with a message."""
    import numpy as np
    from lqrrt_b200 import _native, configs
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    prob = configs.demo_boat_novice(0)
    if not has_gpu:
        with pytest.raises(_native.NativeError) as e:
            _native.Engine(prob.system, 20, 0.1, 0.5, prob.error_tol, 100, 16)
        assert 'no CUDA device' in str(e.value) or 'CUDA' in str(e.value)
    lib = _native.load_library()
    cfg = _native.lq_config()
    cfg.abi_version = 999
    h = ctypes.c_void_p()
    assert lib.lq_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b'ABI' in lib.lq_last_error(None)
    assert lib.lq_set_goal(None, None, None) != 0          # NULL handle is an error, not a crash
