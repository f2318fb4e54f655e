"""User-compiled dynamics (lqrrt_b200/plugin.py) on a CPU-only box: the generated
source compiles for sm_100a against the engine's headers, the library exports the factory, registers with
the main library, and every failure is loud (no CPU fallback)."""
import ctypes
import os

import pytest

from lqrrt_b200 import _native, models, plugin
import plugin_models as pm


@pytest.mark.parametrize('name,dtype,lqr_kind', pm.BUILDS)
def test_plugin_builds_and_registers(name, dtype, lqr_kind):
    dyn = pm.dynamics_of(name)
    path = dyn.build(dtype, lqr_kind)
    assert os.path.isfile(path) and path.startswith(plugin.PLUGIN_DIR)
    assert dyn.build(dtype, lqr_kind) == path                       # cached by content hash
    lib = ctypes.CDLL(path)
    lib.lq_plugin_kind.restype = ctypes.c_int
    assert lib.lq_plugin_kind() == dyn.kind >= plugin.PLUGIN_BASE
    assert lib.lq_plugin_abi_version() == _native.LQ_ABI_VERSION
    assert hasattr(lib, 'lq_plugin_make')
    dyn.prepare(_native.load_library(), dtype, lqr_kind)            # dlopen + lq_register_plugin
    assert (dyn.kind, dtype, int(lqr_kind)) in plugin._loaded


def test_plugin_kind_is_stable_and_content_addressed():
    a = plugin.UserDynamics(4, 2, pm.DINT2_STEP, params=[1.0])
    b = plugin.UserDynamics(4, 2, pm.DINT2_STEP, params=[2.5], name='other')     # parameters are data, not code
    c = plugin.UserDynamics(4, 2, pm.DINT2_STEP + '\n', params=[1.0])
    assert a.kind == b.kind and a.kind != c.kind
    assert a.library_path('f32', models.LQR_CONSTANT) != a.library_path('f64', models.LQR_CONSTANT)


def test_plugin_errors_are_loud():
    with pytest.raises(ValueError):
        plugin.UserDynamics(13, 2, 'xn[0] = x[0];')
    with pytest.raises(ValueError):
        plugin.UserDynamics(2, 1, '')
    bad = plugin.UserDynamics(2, 1, 'xn[0] = undeclared_symbol;', name='bad')
    with pytest.raises(plugin.PluginError) as e:
        bad.build('f32', models.LQR_CONSTANT)
    assert 'undeclared_symbol' in str(e.value)
    with pytest.raises(plugin.PluginError):
        bad.source('f32', models.LQR_PD_ROTATED)                    # vehicle-specific provider
    # an unregistered plugin kind is refused by lq_create (after the device check on a GPU box)
    lib = _native.load_library()
    cfg = _native.lq_config()
    cfg.abi_version = _native.LQ_ABI_VERSION; cfg.n = 2; cfg.m = 1; cfg.sys_kind = plugin.PLUGIN_BASE + 7
    h = ctypes.c_void_p()
    assert lib.lq_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert lib.lq_register_plugin(5, 0, 2, None) != 0               # below LQ_SYS_PLUGIN_BASE / null factory
