"""
ctypes binding of the C ABI in include/lqrrt_b200.h (liblqrrt_b200.so, built
from csrc/ for sm_100a).  This is the only module that touches the shared
library; there is no CPU implementation behind it, so a missing library or a
missing GPU is an error, never a fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import models as M

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liblqrrt_b200.so')

LQ_ABI_VERSION = 2
LQ_MAX_N = 12
LQ_F32, LQ_F64 = 0, 1
LQ_OPT_NEAREST_KERNEL, LQ_OPT_SAMPLE_PREFETCH = 1, 2
LQ_OPT_NEAREST_CTAS_PER_SM = 3
LQ_OPT_SPECULATIVE_SCAN, LQ_OPT_FORCE_RESCAN, LQ_OPT_STAGE_TABLES, LQ_OPT_STEER_BLOCK, LQ_OPT_PDL = 4, 5, 6, 7, 8
LQ_OPT_NEAREST_GRID, LQ_OPT_GRID_MIN_KPAIRS, LQ_OPT_SAMPLE_AHEAD, LQ_OPT_WHITEN_METRIC = 9, 10, 11, 12
NEAREST_KERNELS = {'scalar': 0, 'packed': 1, 'packed_q8': 2, 'two_level': 3, 'prefix': 4}


class NativeError(RuntimeError):
    pass


class lq_feas_prim(C.Structure):
    _fields_ = [('kind', C.c_int32), ('n_obs', C.c_int32), ('n_verts', C.c_int32), ('flags', C.c_int32),
                ('dims', C.c_int32 * 4), ('inflate', C.c_double),
                ('obs', C.POINTER(C.c_double)), ('verts', C.POINTER(C.c_double)),
                ('lo', C.POINTER(C.c_double)), ('hi', C.POINTER(C.c_double))]


class lq_config(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('device', C.c_int32), ('dtype', C.c_int32),
                ('n', C.c_int32), ('m', C.c_int32), ('sys_kind', C.c_int32), ('lqr_kind', C.c_int32),
                ('n_sys_params', C.c_int32), ('n_lqr_params', C.c_int32),
                ('sys_params', C.POINTER(C.c_double)), ('lqr_params', C.POINTER(C.c_double)),
                ('angle_mask', C.c_uint32), ('n_prims', C.c_int32), ('prims', C.POINTER(lq_feas_prim)),
                ('horizon_iters', C.c_int32), ('hfactor', C.c_int32), ('hspan_lo', C.c_int32), ('hspan_hi', C.c_int32),
                ('dt', C.c_double), ('FPR', C.c_double),
                ('error_tol', C.c_double * LQ_MAX_N),
                ('node_capacity', C.c_int32), ('wave_capacity', C.c_int32), ('tries_capacity', C.c_int32)]


class lq_wave_stats(C.Structure):
    _fields_ = [('tree_size', C.c_int32), ('n_sampled', C.c_int32), ('n_accepted', C.c_int32),
                ('n_goal', C.c_int32), ('best_node', C.c_int32), ('best_cost', C.c_int32),
                ('full', C.c_int32), ('tries_used', C.c_int32), ('flags', C.c_int32), ('horizon_iters', C.c_int32),
                ('reserved', C.c_int32 * 2)]


WAVE_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(lq_wave_stats))


class lq_plan_args(C.Structure):
    _fields_ = [('seed', C.c_uint64), ('wave_ctr0', C.c_uint32), ('wave_fixed', C.c_int32),
                ('wave_min', C.c_int32), ('wave_max', C.c_int32), ('growth', C.c_double),
                ('use_ignore', C.c_int32), ('world', C.c_int32), ('rank', C.c_int32), ('depth', C.c_int32),
                ('lag', C.c_int32), ('on_wave', WAVE_CALLBACK), ('user', C.c_void_p)]


class lq_plan_summary(C.Structure):
    _fields_ = [('waves', C.c_int32), ('waves_enqueued', C.c_int32), ('stop_code', C.c_int32), ('reserved', C.c_int32),
                ('samples', C.c_int64), ('device_ms', C.c_double), ('host_enqueue_ms', C.c_double),
                ('host_wait_ms', C.c_double), ('host_callback_ms', C.c_double), ('last', lq_wave_stats)]


# every symbol include/lqrrt_b200.h declares: (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int32)
_U8 = C.POINTER(C.c_uint8)
SYMBOLS = {
    'lq_abi_version': (C.c_int, []),
    'lq_create': (C.c_int, [C.POINTER(lq_config), C.POINTER(_P)]),
    'lq_register_plugin': (C.c_int, [C.c_int, C.c_int, C.c_int, _P]),
    'lq_destroy': (C.c_int, [_P]),
    'lq_last_error': (C.c_char_p, [_P]),
    'lq_set_goal': (C.c_int, [_P, _D, _D]),
    'lq_set_horizon': (C.c_int, [_P, C.c_int, _P]),
    'lq_set_sampling': (C.c_int, [_P, _D, _D, _D, C.c_int]),
    'lq_reset': (C.c_int, [_P, _D, _P]),
    'lq_load_nodes': (C.c_int, [_P, C.c_int, _D, _I, _I, _P]),
    'lq_tree_size': (C.c_int, [_P, C.POINTER(C.c_int)]),
    'lq_nearest': (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P]),
    'lq_steer': (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P]),
    'lq_wave_expand': (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _D, C.c_int, C.c_int, _P]),
    'lq_wave_expand_candidates': (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int, C.c_int, C.POINTER(C.c_int), _P]),
    'lq_wave_records': (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    'lq_wave_commit': (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    'lq_wave_set_peers': (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.c_uint64]),
    'lq_wave_stats_get': (C.c_int, [_P, C.POINTER(lq_wave_stats)]),
    'lq_plan_run': (C.c_int, [_P, C.POINTER(lq_plan_args), C.POINTER(lq_plan_summary), _P]),
    'lq_stop_word': (C.c_int, [_P, C.POINTER(C.POINTER(C.c_int32))]),
    'lq_plan_readout': (C.c_int, [_P, C.c_int, _D, _U8, C.c_int, _I, C.POINTER(C.c_int), _D, _D, _I, _P]),
    'lq_policy_wave': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]),
    'lq_policy_slice': (C.c_int, [C.c_int, C.c_int, C.c_int, _I, _I]),
    'lq_closest': (C.c_int, [_P, _D, _U8, C.POINTER(C.c_int), _P]),
    'lq_steer_force': (C.c_int, [_P, C.c_int, _D, C.c_int, _D, _D, C.POINTER(C.c_int), _P]),
    'lq_append_node': (C.c_int, [_P, C.c_int, _D, _D, _D, C.c_int, _P]),
    'lq_mark_ignore': (C.c_int, [_P, _I, C.c_int, _P]),
    'lq_climb': (C.c_int, [_P, C.c_int, _I, C.c_int, C.POINTER(C.c_int), _P]),
    'lq_download_nodes': (C.c_int, [_P, C.c_int, C.c_int, _D, _I, _I, _I, _U8, _U8, _D, _P]),
    'lq_download_edges': (C.c_int, [_P, _I, C.c_int, _D, _D, _I, _P]),
    'lq_eval_dynamics': (C.c_int, [_P, C.c_int, _D, _D, C.c_double, _D]),
    'lq_eval_lqr': (C.c_int, [_P, C.c_int, _D, _D, _D, _D]),
    'lq_eval_erf': (C.c_int, [_P, C.c_int, _D, _D, _D]),
    'lq_eval_feasible': (C.c_int, [_P, C.c_int, _D, _D, _U8]),
    'lq_launch_count': (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    'lq_set_tuning': (C.c_int, [_P, C.c_int, C.c_int]),
    'lq_set_option': (C.c_int, [_P, C.c_int, C.c_int]),
    'lq_get_option': (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    'lq_set_profiling': (C.c_int, [_P, C.c_int]),
    'lq_profile_read': (C.c_int, [_P, _D, _D, C.c_int]),
}

_lib = None


def load_library():
    """dlopen the in-tree shared library and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeError(
            "{} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()' or make -C lqrrt_b200/csrc). lqrrt_b200 has no CPU fallback.".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)            # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.lq_abi_version() != LQ_ABI_VERSION:
        raise NativeError("liblqrrt_b200.so ABI version {} != binding version {}".format(
            lib.lq_abi_version(), LQ_ABI_VERSION))
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(_D)


def _ip(a):
    return a.ctypes.data_as(_I)


def _c64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


class Engine(object):
    """One lq_engine handle: one GPU's replica of the tree plus the wave
    working set.  Methods map 1:1 onto the C ABI."""

    def __init__(self, system, horizon_iters, dt, FPR, error_tol, node_capacity, wave_capacity,
                 dtype='f32', device=0, tries_capacity=16, hfactor=0, hspan=None):
        self.lib = load_library()
        self.system = system
        self.n = system.nstates; self.m = system.ncontrols
        self.H = int(horizon_iters)
        self.dtype = dtype
        self.np_dtype = np.float32 if dtype == 'f32' else np.float64
        self.device = int(device)
        self.node_capacity = int(node_capacity)
        self.wave_capacity = (int(wave_capacity) + 3) & ~3
        if dtype not in ('f32', 'f64'):
            raise ValueError("dtype must be 'f32' or 'f64'")

        keep = []          # keep numpy buffers alive across lq_create
        cfg = lq_config()
        cfg.abi_version = LQ_ABI_VERSION
        cfg.device = self.device
        cfg.dtype = LQ_F32 if dtype == 'f32' else LQ_F64
        cfg.n = self.n; cfg.m = self.m
        cfg.sys_kind = system.dynamics.kind; cfg.lqr_kind = system.lqr.kind
        sp = _c64(system.dynamics.blob()); lp = _c64(system.lqr.blob(self.n, self.m)); keep += [sp, lp]
        cfg.n_sys_params = len(sp); cfg.n_lqr_params = len(lp)
        cfg.sys_params = _dp(sp); cfg.lqr_params = _dp(lp)
        cfg.angle_mask = system.erf.mask(self.n)
        prims = system.is_feasible.prims
        arr = (lq_feas_prim * max(1, len(prims)))()
        for i, p in enumerate(prims):
            arr[i].kind = p.kind
            arr[i].flags = p.flags
            arr[i].inflate = p.inflate
            for k in range(4):
                arr[i].dims[k] = p.dims[k]
            if p.obs is not None:
                o = _c64(p.obs); keep.append(o)
                arr[i].n_obs = len(o); arr[i].obs = _dp(o)
            if p.verts is not None:
                v = _c64(p.verts); keep.append(v)
                arr[i].n_verts = v.shape[1]; arr[i].verts = _dp(v)
            if p.kind == M.FEAS_OGRID_CLOUD:
                lo = _c64(p.lo); hi = _c64(p.hi); keep += [lo, hi]
                arr[i].lo = _dp(lo); arr[i].hi = _dp(hi)
            elif p.lo is not None:
                want = self.m if p.kind == M.FEAS_EFFORT_BOX else self.n
                lo = _c64(p.lo); hi = _c64(p.hi)
                if len(lo) != want or len(hi) != want:
                    raise ValueError("box primitive needs {} bounds".format(want))
                keep += [lo, hi]
                arr[i].lo = _dp(lo); arr[i].hi = _dp(hi)
        cfg.n_prims = len(prims); cfg.prims = arr
        cfg.horizon_iters = self.H
        cfg.hfactor = int(hfactor)
        cfg.hspan_lo, cfg.hspan_hi = (int(hspan[0]), int(hspan[1])) if hfactor else (self.H, self.H)
        cfg.dt = float(dt); cfg.FPR = float(FPR)
        tol = np.abs(np.asarray(error_tol, dtype=np.float64)) * np.ones(self.n)
        for i in range(self.n):
            cfg.error_tol[i] = tol[i]
        cfg.node_capacity = self.node_capacity
        cfg.wave_capacity = self.wave_capacity
        cfg.tries_capacity = int(tries_capacity)

        prepare = getattr(system.dynamics, 'prepare', None)
        if prepare is not None:         # user-compiled dynamics (plugin.py): build + register its factory first
            prepare(self.lib, dtype, system.lqr.kind)
        h = _P()
        rc = self.lib.lq_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise NativeError("lq_create failed ({}): {}".format(rc, self.lib.lq_last_error(None).decode()))
        self.h = h
        del keep
        self.elt = 4 if dtype == 'f32' else 8
        self.h2d_bytes = 0        # bytes this binding moved host -> device / device -> host
        self.d2h_bytes = 0

    # ------------------------------------------------------------------ utils
    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.lq_last_error(self.h).decode()
            if rc == -1:
                raise ValueError(msg or "invalid argument")
            raise NativeError("lqrrt_b200 native call failed ({}): {}".format(rc, msg))

    def close(self):
        if getattr(self, 'h', None):
            self.lib.lq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream(stream):
        if stream is None:
            return None
        return _P(int(stream))

    # -------------------------------------------------------------- problem
    def set_goal(self, goal, buffer):
        g = _c64(goal); b = _c64(buffer)
        key = g.tobytes() + b.tobytes()
        if key == getattr(self, '_goal_key', None):      # unchanged since the last plan: the device copy stands
            return
        self.h2d_bytes += 2 * LQ_MAX_N * self.elt
        self._ck(self.lib.lq_set_goal(self.h, _dp(g), _dp(b)))
        self._goal_key = key

    def set_horizon(self, horizon_iters, stream=None):
        self._ck(self.lib.lq_set_horizon(self.h, int(horizon_iters), self._stream(stream)))

    def set_sampling(self, sample_space, goal_bias, tries):
        ss = _c64(sample_space, (self.n, 2))
        lo = np.ascontiguousarray(ss[:, 0]); hi = np.ascontiguousarray(ss[:, 1])
        gb = _c64(goal_bias) * np.ones(self.n)
        self._ck(self.lib.lq_set_sampling(self.h, _dp(lo), _dp(hi), _dp(gb), int(tries)))

    def reset(self, x0, stream=None):
        x = _c64(x0)
        self.h2d_bytes += self.n * self.elt + 8
        self._ck(self.lib.lq_reset(self.h, _dp(x), self._stream(stream)))

    def load_nodes(self, state, parent=None, cost=None, stream=None):
        s = _c64(state).reshape(-1, self.n)
        self.h2d_bytes += len(s) * (self.n * self.elt + 8)
        p = None if parent is None else np.ascontiguousarray(parent, dtype=np.int32)
        c = None if cost is None else np.ascontiguousarray(cost, dtype=np.int32)
        self._ck(self.lib.lq_load_nodes(self.h, len(s), _dp(s), None if p is None else _ip(p),
                                        None if c is None else _ip(c), self._stream(stream)))

    def tree_size(self):
        out = C.c_int(0)
        self._ck(self.lib.lq_tree_size(self.h, C.byref(out)))
        return out.value

    # ------------------------------------------------------------ hot path
    def nearest(self, d_queries, W, use_ignore, d_idx, d_cost=None, stream=None):
        """Device pointers (ints): queries n x W dim-major engine dtype; idx W int32."""
        self._ck(self.lib.lq_nearest(self.h, _P(d_queries), int(W), int(bool(use_ignore)), _P(d_idx),
                                     _P(d_cost) if d_cost else None, self._stream(stream)))

    def steer(self, d_ids, d_targets, W, d_xseq=None, d_useq=None, d_len=None, stream=None):
        self._ck(self.lib.lq_steer(self.h, _P(d_ids), _P(d_targets), int(W),
                                   _P(d_xseq) if d_xseq else None, _P(d_useq) if d_useq else None,
                                   _P(d_len) if d_len else None, self._stream(stream)))

    def wave_expand(self, seed, wave_ctr, first, count, samples=None, use_ignore=True, flags=0, stream=None):
        hs = None
        if samples is not None:
            hs = _c64(samples).reshape(-1, self.n)
            if len(hs) != count:
                raise ValueError("samples must be count x n")
            self.h2d_bytes += hs.size * self.elt
        self._ck(self.lib.lq_wave_expand(self.h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_uint32(int(wave_ctr) & 0xFFFFFFFF),
                                         int(first), int(count), None if hs is None else _dp(hs),
                                         int(bool(use_ignore)), int(flags), self._stream(stream)))

    def wave_expand_candidates(self, count, tries, candidates, use_ignore=True, stream=None):
        c = _c64(candidates).reshape(-1, self.n)
        self.h2d_bytes += c.size * self.elt; self.d2h_bytes += 4
        used = C.c_int(0)
        self._ck(self.lib.lq_wave_expand_candidates(self.h, int(count), int(tries), _dp(c), len(c),
                                                    int(bool(use_ignore)), C.byref(used), self._stream(stream)))
        return used.value

    def wave_records(self):
        p = _P(); b = C.c_uint64(0)
        self._ck(self.lib.lq_wave_records(self.h, C.byref(p), C.byref(b)))
        return p.value, b.value

    def wave_commit(self, d_records=None, n_segments=1, stream=None, my_segment=0):
        self._ck(self.lib.lq_wave_commit(self.h, _P(d_records) if d_records else None, int(n_segments),
                                         int(my_segment), self._stream(stream)))

    def set_peers(self, world, rank, ptrs, nbytes):
        arr = (_P * max(1, len(ptrs)))(*[_P(int(p)) for p in ptrs])
        self._ck(self.lib.lq_wave_set_peers(self.h, int(world), int(rank), arr, C.c_uint64(int(nbytes))))

    def wave_stats(self):
        st = lq_wave_stats()
        self.d2h_bytes += C.sizeof(lq_wave_stats)
        self._ck(self.lib.lq_wave_stats_get(self.h, C.byref(st)))
        return st

    def plan_run(self, seed, wave_ctr0, wave_fixed, wave_min, wave_max, growth, use_ignore, on_wave=None,
                 world=1, rank=0, depth=0, stream=None, lag=1):
        """The whole expand loop of one plan on the device-scheduled pipeline (lq_plan_run).  `on_wave(stats)`
        is called once per committed wave and returns 0 / 1 (done) / 2 (kill).  Returns the summary struct."""
        err = []

        def _cb(_user, st_ptr):
            try:
                return int(on_wave(st_ptr.contents)) if on_wave is not None else 0
            except BaseException as e:          # never unwind through the C frame: stop the plan, re-raise after
                err.append(e)
                return 2
        cb = WAVE_CALLBACK(_cb)
        a = lq_plan_args()
        a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF; a.wave_ctr0 = int(wave_ctr0) & 0xFFFFFFFF
        a.wave_fixed = int(wave_fixed or 0); a.wave_min = int(wave_min); a.wave_max = int(wave_max)
        a.growth = float(growth); a.use_ignore = int(bool(use_ignore)); a.world = int(world); a.rank = int(rank)
        a.depth = int(depth); a.lag = int(lag); a.on_wave = cb; a.user = None
        out = lq_plan_summary()
        rc = self.lib.lq_plan_run(self.h, C.byref(a), C.byref(out), self._stream(stream))
        if err:
            raise err[0]
        self._ck(rc)
        self.d2h_bytes += out.waves * 64
        return out

    def plan_readout(self, node=None, xguide=None, colmask=None, max_len=4096, stream=None):
        """[closest-to-guide ->] climb -> edges of the path in one call.  Returns (ids, hx, hu, edge_len); grows
        max_len and retries if the path is longer."""
        xg = None if xguide is None else _c64(xguide)
        mk = None if colmask is None else np.ascontiguousarray(colmask, dtype=np.uint8)
        while True:
            # the engine fills the first `len` rows; nothing else is read (no zero-fill of max_len rows per plan)
            ids = np.empty(max_len, dtype=np.int32); ln = C.c_int(0)
            hx = np.empty((max_len, self.H, self.n)); hu = np.empty((max_len, self.H, self.m))
            el = np.empty(max_len, dtype=np.int32)
            rc = self.lib.lq_plan_readout(self.h, -1 if node is None else int(node), None if xg is None else _dp(xg),
                                          None if mk is None else mk.ctypes.data_as(_U8), int(max_len), _ip(ids), C.byref(ln),
                                          _dp(hx), _dp(hu), _ip(el), self._stream(stream))
            if rc == -3 and b'longer than' in self.lib.lq_last_error(self.h) and max_len < (1 << 22):
                max_len *= 8
                continue
            self._ck(rc)
            break
        L = ln.value
        self.h2d_bytes += self.n * (self.n + 1) * self.elt if node is None else 4
        self.d2h_bytes += 4 * (max_len + 2) + L * (self.H * (self.n + self.m) * self.elt + 4)
        return ids[:L].tolist(), hx[:L], hu[:L], el[:L]

    def request_stop(self, code):
        """Write the stop word the rollout kernels poll (callable from any thread; 1 = done, 2 = kill)."""
        if getattr(self, '_stop_ptr', None) is None:
            p = C.POINTER(C.c_int32)()
            self._ck(self.lib.lq_stop_word(self.h, C.byref(p)))
            self._stop_ptr = p
        self._stop_ptr[0] = int(code)

    def closest(self, xguide, colmask, stream=None):
        x = _c64(xguide)
        self.h2d_bytes += self.n * (self.n + 1) * self.elt; self.d2h_bytes += 4
        mk = np.ascontiguousarray(colmask, dtype=np.uint8)
        out = C.c_int(0)
        self._ck(self.lib.lq_closest(self.h, _dp(x), mk.ctypes.data_as(_U8), C.byref(out), self._stream(stream)))
        return out.value

    def steer_force(self, node, xtar, max_steps, stream=None):
        xt = _c64(xtar)
        hx = np.zeros((max_steps, self.n)); hu = np.zeros((max_steps, self.m))
        ln = C.c_int(0)
        self._ck(self.lib.lq_steer_force(self.h, int(node), _dp(xt), int(max_steps), _dp(hx), _dp(hu), C.byref(ln),
                                         self._stream(stream)))
        return hx[:ln.value].copy(), hu[:ln.value].copy()

    def append_node(self, parent, state, x_seq, u_seq, stream=None):
        s = _c64(state)
        x = _c64(x_seq).reshape(-1, self.n); u = _c64(u_seq).reshape(-1, self.m)
        self._ck(self.lib.lq_append_node(self.h, int(parent), _dp(s), _dp(x), _dp(u), len(x), self._stream(stream)))

    def mark_ignore(self, ids, stream=None):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        self._ck(self.lib.lq_mark_ignore(self.h, _ip(a), len(a), self._stream(stream)))

    def climb(self, node, max_len=1 << 16, stream=None):
        ids = np.zeros(max_len, dtype=np.int32); ln = C.c_int(0)
        self._ck(self.lib.lq_climb(self.h, int(node), _ip(ids), int(max_len), C.byref(ln), self._stream(stream)))
        self.d2h_bytes += 4 * ln.value + 4
        return ids[:ln.value].tolist()

    # -------------------------------------------------------------- read-back
    def download_nodes(self, first, count, want_K=False, stream=None):
        state = np.zeros((count, self.n)); parent = np.zeros(count, dtype=np.int32)
        cost = np.zeros(count, dtype=np.int32); elen = np.zeros(count, dtype=np.int32)
        ign = np.zeros(count, dtype=np.uint8); own = np.zeros(count, dtype=np.uint8)
        K = np.zeros((count, self.m, self.n)) if want_K else None
        self.d2h_bytes += count * (self.n * self.elt + 14 + (self.m * self.n * self.elt if want_K else 0))
        self._ck(self.lib.lq_download_nodes(self.h, int(first), int(count), _dp(state), _ip(parent), _ip(cost), _ip(elen),
                                            ign.ctypes.data_as(_U8), own.ctypes.data_as(_U8),
                                            None if K is None else _dp(K), self._stream(stream)))
        return dict(state=state, parent=parent, cost=cost, edge_len=elen, ignore=ign, owner=own, K=K)

    def download_edges(self, ids, stream=None):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        hx = np.zeros((len(a), self.H, self.n)); hu = np.zeros((len(a), self.H, self.m))
        self.h2d_bytes += 4 * len(a); self.d2h_bytes += len(a) * self.H * (self.n + self.m) * self.elt
        hl = np.zeros(len(a), dtype=np.int32)
        if len(a):
            self._ck(self.lib.lq_download_edges(self.h, _ip(a), len(a), _dp(hx), _dp(hu), _ip(hl), self._stream(stream)))
        return hx, hu, hl

    # -------------------------------------------------------------- callbacks
    def eval_dynamics(self, x, u, dt):
        x = _c64(x).reshape(-1, self.n); u = _c64(u).reshape(-1, self.m)
        out = np.zeros_like(x)
        self._ck(self.lib.lq_eval_dynamics(self.h, len(x), _dp(x), _dp(u), float(dt), _dp(out)))
        return out

    def eval_lqr(self, x, u):
        x = _c64(x).reshape(-1, self.n); u = _c64(u).reshape(-1, self.m)
        S = np.zeros((len(x), self.n, self.n)); K = np.zeros((len(x), self.m, self.n))
        self._ck(self.lib.lq_eval_lqr(self.h, len(x), _dp(x), _dp(u), _dp(S), _dp(K)))
        return S, K

    def eval_erf(self, xgoal, x):
        xg = _c64(xgoal).reshape(-1, self.n); x = _c64(x).reshape(-1, self.n)
        out = np.zeros_like(x)
        self._ck(self.lib.lq_eval_erf(self.h, len(x), _dp(xg), _dp(x), _dp(out)))
        return out

    def eval_feasible(self, x, u):
        x = _c64(x).reshape(-1, self.n); u = _c64(u).reshape(-1, self.m)
        ok = np.zeros(len(x), dtype=np.uint8)
        self._ck(self.lib.lq_eval_feasible(self.h, len(x), _dp(x), _dp(u), ok.ctypes.data_as(_U8)))
        return ok.astype(bool)

    def launch_count(self):
        out = C.c_uint64(0)
        self._ck(self.lib.lq_launch_count(self.h, C.byref(out)))
        return out.value

    def set_tuning(self, steer_lanes=0, nearest_q=0):
        self._ck(self.lib.lq_set_tuning(self.h, int(steer_lanes), int(nearest_q)))

    def set_option(self, option, value):
        self._ck(self.lib.lq_set_option(self.h, int(option), int(value)))

    def get_option(self, option):
        out = C.c_int(0)
        self._ck(self.lib.lq_get_option(self.h, int(option), C.byref(out)))
        return out.value

    def set_profiling(self, on=True):
        self._ck(self.lib.lq_set_profiling(self.h, int(bool(on))))

    def profile_read(self, reset=False):
        ms = np.zeros(5); work = np.zeros(4)
        self._ck(self.lib.lq_profile_read(self.h, _dp(ms), _dp(work), int(bool(reset))))
        return dict(ms=dict(zip(('sample', 'nearest_old', 'nearest_delta', 'steer', 'commit'), ms.tolist())),
                    pairs=work[0], nearest_bytes=work[1], samples=work[2], waves=work[3])


class ModelEvaluator(object):
    """Backs the bare-call form of the specs (`sys.dynamics(x, u, dt)` ...):
    a minimal float64 engine used only for its lq_eval_* entry points."""

    def __init__(self, system):
        self.system = system
        self.engine = Engine(system, horizon_iters=1, dt=system.dt_hint, FPR=0.0,
                             error_tol=np.zeros(system.nstates), node_capacity=4, wave_capacity=4,
                             dtype='f64', device=_current_device())

    def dynamics(self, x, u, dt):
        return self.engine.eval_dynamics(x, np.atleast_1d(u), dt)[0]

    def lqr(self, x, u):
        u = np.zeros(self.system.ncontrols) + np.asarray(u, dtype=np.float64)
        S, K = self.engine.eval_lqr(x, u)
        return S[0], K[0]

    def erf(self, xgoal, x):
        return self.engine.eval_erf(xgoal, x)[0]

    def is_feasible(self, x, u):
        u = np.zeros(self.system.ncontrols) + np.asarray(u, dtype=np.float64)
        return bool(self.engine.eval_feasible(x, u)[0])


def _current_device():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


class DevicePointer(object):
    """Exposes a raw device allocation through __cuda_array_interface__ so torch can wrap it without a
    copy (torch.as_tensor): the engine's record buffer becomes the allgather input in place."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}
