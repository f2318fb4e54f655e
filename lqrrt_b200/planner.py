"""
Planner: the reference's public planning class (lqrrt/planner.py) with its
expand loop moved onto the GPU.

Same constructor, setters, result attributes and error messages as the
reference (planner.py:85-100, 468-610); `update_plan` keeps the reference's
stop rules and close-out paths (planner.py:286-336) but each pass of its loop
expands a *wave* of samples through the C ABI (sample -> nearest -> steer ->
commit, include/lqrrt_b200.h) instead of one sample in Python.

New keyword arguments (all optional, defaults keep reference call sites valid):
  wave      None = adaptive wave size, int = fixed global wave size
            (1 reproduces the reference's sequential semantics)
  wave_max  cap of the adaptive wave size;  wave_growth: adaptive W = wave_growth * tree size (default 2: what bench.py
            measures; 1 spends ~30 more small waves on a 100k-node tree for a deeper tree per sample)
  dtype     'f32' (throughput) or 'f64' (reference-parity arithmetic)
  rng       'philox' = counter-based device sampler; 'numpy' = the reference's
            global MT19937 stream, drawn on the host in the reference's order
  seed      Philox key
  device    CUDA ordinal (default: torch's current device)
  pipeline  enqueue wave k+1 before reading wave k's counters (device sampler only; default on)
  exchange  'peer' = fused pack + exchange by NVLink peer stores into symmetric memory (default),
            'nccl' = one all_gather_into_tensor per wave
  nearest_kernel  'scalar' | 'packed' (FP32x2 issue + two-level argmin; float32 diagonal metrics only);
            None = $LQRRT_NEAREST or DEFAULT_NEAREST_KERNEL.  Same tree either way.
  prefetch  issue each wave's device sampler on a side stream so it overlaps the previous wave
            (needs pipeline); None = $LQRRT_PREFETCH or DEFAULT_PREFETCH.  Same tree either way.
  comm      torch.distributed process group (or True for WORLD): the wave is
            sharded over the ranks and accepted nodes are all-gathered so every
            rank's tree stays identical
  schedule  'device' (default where it applies: device sampler, pipeline on, wave != 1, exchange 'peer' or one
            GPU) = the whole loop runs as lq_plan_run: waves are sized on the device from the tree size the
            previous wave started with, enqueued ahead of the host, and the scan of the old part of the tree
            overlaps the previous wave's rollouts; 'host' = one lq_wave_expand / lq_wave_commit per wave, sized
            by the host.  Both grow the same tree for the same wave sizes.
  lag       device schedule only.  1 (default): wave k searches the tree as committed by wave k-1.
            2: pipelined waves -- wave k searches the tree as committed by wave k-2, so two consecutive waves'
            scans and rollouts are in flight together and only the commits are serial.  Waves complete faster once
            they are latency-bound (several GPUs), but the tree gets one edge deeper per two waves instead of per wave:
            the choice for filling a tree fast (bench.py's throughput leg), not for reaching a distant goal with
            small waves (measured on C4: lag 1 reaches the goal 283 m away in 56 ms, lag 2 fills 100k nodes first).
            oracle/wave.py defines both.
  fresh_samples  True (default): the Philox wave counter keeps counting across update_plan calls, so a replan
            draws new samples as the reference's global NumPy stream does; assigning `planner.seed` restarts
            it.  False: every plan replays the same sample sequence.

dynamics / lqr / erf / constraints.is_feasible must be `lqrrt_b200.models`
specs; there is no CPU fallback for arbitrary Python callables.
"""
from __future__ import annotations

import os
import time

import numpy as np
import scipy.interpolate

from . import models as _models
from . import waves as _waves
from .constraints import Constraints
from .tree import Tree


# Issue-level options of the engine (include/lqrrt_b200.h lq_set_option); results do not depend on them.
# Measured on B200 (profiles/README.md, round 1): packed scan +3 % (C4) / +11 % (C5k) on the scan's time,
# prefetch takes the sampler (1.0 of 14.7 ms per C4 plan) off the wave's dependent chain.  Prefetch is on by
# default for single-GPU planners only: it is verified bit-identical there (tests/test_gpu_zz_options.py);
# multi-GPU planners take it when asked (prefetch=True or $LQRRT_PREFETCH=1).
DEFAULT_NEAREST_KERNEL = 'packed'
DEFAULT_PREFETCH = True


class Planner(object):

    def __init__(self, dynamics, lqr, constraints,
                 horizon, dt=0.05, FPR=0,
                 error_tol=0.05, erf=np.subtract,
                 min_time=0.5, max_time=1, max_nodes=1E5,
                 goal0=None, sys_time=time.time, printing=True,
                 wave=None, wave_max=8192, dtype='f32', rng='philox', seed=0, device=None, comm=None,
                 wave_growth=2.0, pipeline=True, exchange='peer', nearest_kernel=None, prefetch=None,
                 schedule=None, fresh_samples=True, depth=0, lag=None):
        self._engine = None
        if schedule not in (None, 'device', 'host'):
            raise ValueError("schedule must be 'device' or 'host'")
        self.schedule = schedule or os.environ.get('LQRRT_SCHEDULE') or 'device'
        self.fresh_samples = bool(fresh_samples)
        self.depth = int(depth)
        if lag is None:
            lag = int(os.environ.get('LQRRT_LAG') or 1)
        if lag not in (1, 2):
            raise ValueError("lag must be 1 or 2")
        self.lag = lag
        self._wave_base = 0
        self.engine_options = {}       # {LQ_OPT_*: value} applied to the engine before every plan (tests, tuning)
        if nearest_kernel is None:
            nearest_kernel = os.environ.get('LQRRT_NEAREST') or DEFAULT_NEAREST_KERNEL
        from ._native import NEAREST_KERNELS
        if nearest_kernel not in NEAREST_KERNELS:
            raise ValueError("nearest_kernel must be one of {}".format(sorted(NEAREST_KERNELS)))
        self.nearest_kernel = nearest_kernel
        self._prefetch_explicit = True
        if prefetch is None:
            env = os.environ.get('LQRRT_PREFETCH')
            self._prefetch_explicit = env not in (None, '')
            prefetch = DEFAULT_PREFETCH if not self._prefetch_explicit else env not in ('0', 'false', 'off')
        self.prefetch = bool(prefetch)
        if exchange not in ('peer', 'nccl'):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange_mode = exchange
        self.pipeline = bool(pipeline)
        self._engine_key = None
        self._exchange = None
        self.wave = wave
        self.wave_max = int(wave_max)
        self.wave_growth = float(wave_growth)
        self.dtype = dtype
        if rng not in ('philox', 'numpy'):
            raise ValueError("rng must be 'philox' or 'numpy'")
        self.rng = rng
        self.seed = int(seed)
        self.device = device
        self._set_comm(comm)

        self.set_system(dynamics, lqr, constraints, erf)
        self.set_resolution(horizon, dt, FPR, error_tol)
        self.set_runtime(min_time, max_time, max_nodes, sys_time)
        self.set_goal(goal0)
        self.printing = printing
        self.killed = False
        self.stats = {}

    @property
    def seed(self):
        return self._seed

    @seed.setter
    def seed(self, value):
        self._seed = int(value)
        self._wave_base = 0           # a newly assigned key starts its sample stream from the beginning

    # ------------------------------------------------------------------ comm
    def _set_comm(self, comm):
        self.comm = None
        self.world = 1
        self.rank = 0
        if comm is None or comm is False:
            return
        import torch.distributed as dist
        if comm is True:
            comm = dist.group.WORLD
        self.comm = comm
        self.world = dist.get_world_size(comm)
        self.rank = dist.get_rank(comm)

    @property
    def prefetch_active(self):
        return bool(self.prefetch and self.pipeline)

    # ---------------------------------------------------------------- engine
    def _device(self):
        if self.device is not None:
            return int(self.device)
        from ._native import _current_device
        return _current_device()

    def _wave_capacity(self):
        top = self.wave_max if self.wave is None else max(1, int(self.wave))
        return _waves.local_capacity(top, self.world)

    def _ensure_engine(self):
        """(Re)build the device engine when anything baked into it changed; `dt` and friends are plain
        attributes in the reference (lqrrt_node.py:259 assigns planner.dt directly), so compare by value."""
        from ._native import Engine
        # constraints.set_feasibility_function may have swapped the spec since set_system
        self._system = _models.bundle_of(self.dynamics, self.lqr, self.erf, self.constraints.is_feasible,
                                         self.nstates, self.ncontrols)
        tol = np.abs(np.asarray(self.error_tol, dtype=np.float64)) * np.ones(self.nstates)
        # with the adaptive heuristic horizon_iters is run-time state (planner.py:421,424) and the longest edge
        # is hspan[1]; with a fixed horizon it is baked in
        H = int(self.hspan[1]) if self.hfactor else int(self.horizon_iters)
        key = (_models.system_digest(self._system), H, int(self.hfactor), int(self.hspan[0]), int(self.hspan[1]), float(self.dt),
               float(self.FPR), tuple(tol.tolist()),
               int(self.max_nodes) + 1, self._wave_capacity(), self.dtype, self._device())
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self._system, H, self.dt, self.FPR, tol,
                                  node_capacity=int(self.max_nodes) + 1, wave_capacity=self._wave_capacity(),
                                  dtype=self.dtype, device=self._device(), hfactor=int(self.hfactor),
                                  hspan=(int(self.hspan[0]), int(self.hspan[1])))
            self._engine_key = key
            self._exchange = None
            tune = os.environ.get('LQRRT_TUNING')          # "steer_lanes[,nearest_q]" for experiments
            if tune:
                t = [int(v) for v in tune.split(',')] + [0]
                self._engine.set_tuning(t[0], t[1])
        from ._native import LQ_OPT_NEAREST_KERNEL, LQ_OPT_SAMPLE_PREFETCH, LQ_OPT_NEAREST_CTAS_PER_SM, NEAREST_KERNELS
        self._engine.set_option(LQ_OPT_NEAREST_KERNEL, NEAREST_KERNELS[self.nearest_kernel])
        self._engine.set_option(LQ_OPT_SAMPLE_PREFETCH, int(self.prefetch_active))
        for env, opt in (('LQRRT_SPEC', 4), ('LQRRT_STEER_BLOCK', 7), ('LQRRT_SCAN_CTAS', LQ_OPT_NEAREST_CTAS_PER_SM), ('LQRRT_STAGE', 6), ('LQRRT_PDL', 8), ('LQRRT_GRID', 9), ('LQRRT_GRID_MIN', 10), ('LQRRT_SAMPLE_AHEAD', 11), ('LQRRT_WHITEN', 12)):
            if os.environ.get(env):                          # experiments (tools/gpu_*.sh): issue-level options only
                self._engine.set_option(opt, int(os.environ[env]))
        for opt, val in self.engine_options.items():
            self._engine.set_option(opt, val)
        return self._engine

    @staticmethod
    def _cuda_stream():
        try:
            import torch
            if torch.cuda.is_available():
                return torch.cuda.current_stream().cuda_stream
        except ImportError:
            pass
        return None

    # ----------------------------------------------------------- update_plan
    def update_plan(self, x0, sample_space, goal_bias=0,
                    guide=None, xrand_gen=None, pruning=True,
                    finish_on_goal=False, specific_time=None):
        """
        Grows a new tree from x0 toward the goal and prepares get_state(t) /
        get_effort(t); arguments and return value as reference planner.py:104-154.
        Returns False if it was killed, if the tree exceeded max_nodes, or if no
        goal has been set.
        """
        x0 = np.array(x0, dtype=np.float64)
        if self.goal is None:
            print("No goal has been set yet!")
            self.get_state = lambda t: x0
            self.get_effort = lambda t: np.zeros(self.ncontrols)
            return False

        if specific_time is None:
            min_time, max_time = self.min_time, self.max_time
        else:
            min_time = max_time = specific_time

        # ---- sampler (planner.py:176-216)
        mode = None
        tries_limit = 10
        if xrand_gen is None or type(xrand_gen) is int:
            if goal_bias is None:
                goal_bias = [0] * self.nstates
            elif hasattr(goal_bias, '__contains__'):
                if len(goal_bias) != self.nstates:
                    raise ValueError("Expected goal_bias to be scalar or have same length as state.")
            else:
                goal_bias = [goal_bias] * self.nstates
            # the reference leans on Python 2's `None > 0 == False` here (planner.py:188-191)
            if xrand_gen is not None and xrand_gen > 0:
                tries_limit = int(xrand_gen)
            sample_space = np.array(sample_space, dtype=np.float64)
            if sample_space.shape != (self.nstates, 2):
                raise ValueError("Expected sample_space to be list of nstates tuples.")
            mode = self.rng
            if mode == 'numpy' and self.world > 1:
                raise ValueError("rng='numpy' replays one serial MT19937 stream and cannot be sharded; use rng='philox'")
        else:
            if not hasattr(xrand_gen, '__call__'):
                raise ValueError("Expected xrand_gen to be None, an integer >= 1,  or a function.")
            mode = 'callable'

        t_call = time.perf_counter()
        eng = self._ensure_engine()
        stream = self._cuda_stream()
        eng.set_goal(self.goal, self.constraints.goal_buffer)
        if mode in ('philox', 'numpy'):
            if tries_limit > 16:
                raise ValueError("at most 16 sampler tries are supported")
            eng.set_sampling(sample_space, goal_bias, tries_limit)
            centers = np.mean(sample_space, axis=1)
            spans = np.diff(sample_space).flatten()
            bias_arr = np.asarray(goal_bias, dtype=np.float64)

        if self.hfactor:
            # planner-level state carried across steers and plans (planner.py:421,424,545)
            eng.set_horizon(int(np.clip(self.horizon_iters, self.hspan[0], self.hspan[1])), stream)

        # ---- new tree (planner.py:172-173)
        eng.reset(x0, stream)
        self.tree = Tree(eng, x0, self._system, comm=self.comm, rank=self.rank, world=self.world)

        self.xguide = np.copy(self.goal) if guide is None else np.array(guide, dtype=np.float64)

        if self.printing:
            print("\n...planning...")
        self.plan_reached_goal = False
        self.T = np.inf
        time_elapsed = 0
        best_node = -1
        wave_ctr = 0
        samples_total = 0
        stop_request = _waves.STOP_NONE
        outcome = None
        exchange = self._get_exchange(eng) if self.world > 1 else None
        launches0 = eng.launch_count()
        h2d0, d2h0 = eng.h2d_bytes, eng.d2h_bytes
        ev0 = ev1 = None
        if getattr(self, 'profile', False):
            import torch
            eng.set_profiling(True)
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        time_start = self.sys_time()
        t_loop0 = time.perf_counter()

        # Device-scheduled loop (lq_plan_run): nothing a wave needs comes from the host, so the engine sizes,
        # samples, scans, rolls out and commits wave after wave by itself; the host only reads each wave's
        # counters, applies the reference's stop rules on its own clock and raises a stop word.
        use_device = (self.schedule == 'device' and mode == 'philox' and self.pipeline and self.wave != 1 and
                      (self.world == 1 or isinstance(exchange, _waves.PeerExchange)))
        if use_device:
            state = dict(best_node=-1, elapsed=0.0, waves=0, samples=0)

            def on_wave(st):
                if 'stop' not in state:
                    state['waves'] += 1          # waves that expanded samples (the one that saw the stop included)
                state['samples'] += st.n_sampled
                self.tree._set_size(st.tree_size)
                if self.hfactor:
                    self.horizon_iters = int(st.horizon_iters)
                if st.best_node >= 0 and st.best_node != state['best_node']:        # planner.py:260-283
                    state['best_node'] = st.best_node
                    self.plan_reached_goal = True
                    self.T = (1 + st.best_cost) * self.dt
                    if self.printing:
                        print("Found plan at elapsed time: {} s".format(np.round(state['elapsed'], 6)))
                if st.flags:                     # the device has already stopped: these are the no-op tail waves
                    if 'stop' not in state:
                        state['stop'] = st.flags
                    return 0
                state['elapsed'] = self.sys_time() - time_start
                if self.killed:                                                        # planner.py:289
                    return _waves.STOP_KILL
                if (self.plan_reached_goal and state['elapsed'] >= min_time) or \
                        state['elapsed'] >= max_time or self.tree.size > self.max_nodes:    # planner.py:293, 311
                    return _waves.STOP_DONE
                return _waves.STOP_NONE

            summ = eng.plan_run(self.seed, self._wave_base, 0 if self.wave is None else int(self.wave), 32, self.wave_max,
                                self.wave_growth, pruning, on_wave, world=self.world, rank=self.rank, depth=self.depth,
                                stream=stream, lag=self.lag)
            wave_ctr = state['waves']
            waves_enqueued = summ.waves_enqueued
            samples_total = int(summ.samples)
            time_elapsed = state['elapsed']
            best_node = state['best_node']
            self.tree._set_size(summ.last.tree_size)
            stop = state.get('stop', summ.stop_code)
            # a full tree stops the device with code 1 as well; `killed` is what tells a kill from a normal end
            outcome = 'killed' if (stop == _waves.STOP_KILL or (self.killed and stop != 0)) else 'done'
            self.stats_device_ms = float(summ.device_ms)
            self.stats_host = dict(enqueue_ms=float(summ.host_enqueue_ms), wait_ms=float(summ.host_wait_ms),
                                   callback_ms=float(summ.host_callback_ms))

        # Software pipeline: with the device sampler nothing a wave needs comes from the host, so wave k+1
        # is enqueued before the host blocks on wave k's counters and the GPU never waits for the round
        # trip.  Host-fed samplers (the reference's MT19937 stream, user callables that may look at the
        # tree) stay strictly synchronous, as does wave=1.
        # Wave k is sized from the tree size wave k-1 started with (waves 0 and 1 from the initial tree): that is
        # all the host can know when it runs one wave ahead, it does not depend on the pipeline depth, and it is
        # the rule the device-scheduled loop follows, so both drivers grow the same tree.
        depth = 1 if (mode == 'philox' and self.pipeline and self.wave != 1) else 0
        inflight = 0
        stopping = False
        sizes = [1]                      # sizes[k] = tree size wave k starts from

        def launch(W):
            first, count = _waves.partition(W, self.rank, self.world)
            if mode == 'philox':
                eng.wave_expand(self.seed, self._wave_base + wave_ctr, first, count, None, pruning, stop_request, stream)
            elif mode == 'numpy':
                # planner.py:201-211 on the reference's own stream: draw the most tries this wave could
                # use from a snapshot, let the device pick, then advance the real stream by what was used
                snap = np.random.get_state()
                U = np.random.sample(count * tries_limit * (self.nstates + 1)).reshape(-1, self.nstates + 1)
                cand = centers + spans * (U[:, :self.nstates] - 0.5)
                pick = np.greater(bias_arr[None, :], U[:, self.nstates:])
                cand = np.where(pick, self.goal[None, :], cand)
                used = eng.wave_expand_candidates(count, tries_limit, cand, pruning, stream)
                np.random.set_state(snap)
                np.random.sample(used * (self.nstates + 1))
            else:
                xs = np.array([np.array(xrand_gen(self), dtype=np.float64) for _ in range(count)]).reshape(count, self.nstates)
                eng.wave_expand(self.seed, wave_ctr, first, count, xs, pruning, stop_request, stream)
            if exchange is None:
                eng.wave_commit(None, 1, stream)
            elif isinstance(exchange, _waves.PeerExchange):
                eng.wave_commit(None, self.world, stream, my_segment=self.rank)     # records arrived by peer stores
            else:
                recv = exchange.gather(self._send_view(eng))
                eng.wave_commit(recv.data_ptr(), self.world, stream, my_segment=self.rank)

        while not use_device:
            while inflight <= depth and not stopping and (wave_ctr < 2 or len(sizes) >= wave_ctr):
                launch(_waves.wave_size(sizes[max(0, wave_ctr - 1)], self.wave, wave_max=self.wave_max, growth=self.wave_growth))
                wave_ctr += 1
                inflight += 1
            st = eng.wave_stats()
            inflight -= 1
            samples_total += st.n_sampled
            sizes.append(st.tree_size)
            self.tree._set_size(st.tree_size)
            if self.hfactor:
                self.horizon_iters = int(st.horizon_iters)

            # planner.py:260-283: a strictly faster plan replaces the current one
            if st.best_node >= 0 and st.best_node != best_node:
                best_node = st.best_node
                self.plan_reached_goal = True
                self.T = (1 + st.best_cost) * self.dt
                if self.printing:
                    print("Found plan at elapsed time: {} s".format(np.round(time_elapsed, 6)))

            time_elapsed = self.sys_time() - time_start

            # planner.py:286-311, in the reference's order: kill, reached-goal, out of time/nodes
            if self.killed:
                decision = _waves.STOP_KILL
            elif (self.plan_reached_goal and time_elapsed >= min_time) or \
                    time_elapsed >= max_time or self.tree.size > self.max_nodes:
                decision = _waves.STOP_DONE
            else:
                decision = _waves.STOP_NONE
            if self.world > 1:
                # rank 0's decision rides in the header of the next wave it launches; every rank acts on
                # the flag it reads back, so all ranks stop launching after the same wave
                act = st.flags
                if self.rank == 0 and decision != _waves.STOP_NONE:
                    stop_request = decision
            else:
                act = decision
            if act != _waves.STOP_NONE and outcome is None:
                stopping = True
                outcome = 'killed' if act == _waves.STOP_KILL else 'done'
            if stopping and inflight == 0:
                break
        if outcome == 'done':
            outcome = 'reached' if self.plan_reached_goal else 'timeout'

        t_loop1 = time.perf_counter()
        self.stats = dict(waves=wave_ctr, waves_enqueued=waves_enqueued if use_device else wave_ctr,
                          samples=samples_total, nodes=self.tree.size,
                          elapsed=time_elapsed, gpu_launches=eng.launch_count() - launches0,
                          host_setup_s=t_loop0 - t_call, host_loop_s=t_loop1 - t_loop0)
        if ev0 is not None:
            ev1.record(); ev1.synchronize()
            self.stats['device_ms'] = ev0.elapsed_time(ev1)      # wave loop only, on the launching stream
        if use_device:
            self.stats['device_ms'] = self.stats_device_ms       # CUDA events on the pipeline's own chain stream
            self.stats['host_pipeline_ms'] = self.stats_host
        self.stats['schedule'] = 'device' if use_device else 'host'
        self.stats['lag'] = self.lag if use_device else 1
        if self.fresh_samples:
            self._wave_base += waves_enqueued if use_device else wave_ctr

        if self.plan_reached_goal:
            # the plan attributes the reference fills inside its loop (planner.py:276-281)
            self.node_seq, self.x_seq, self.u_seq = self.tree.plan(best_node)
            self.T = len(self.x_seq) * self.dt
            self.t_seq = np.arange(len(self.x_seq)) * self.dt

        if outcome == 'reached':
            if finish_on_goal:                                      # planner.py:294-303
                budget = float(np.clip(min_time / 2, 0.1, np.inf))
                max_steps = int(min(200000, max(1, budget * 10000)))
                xg, ug = eng.steer_force(self.node_seq[-1], self.goal, max_steps, stream)
                if len(xg) >= max_steps and self.printing:
                    print("(exact goal-convergence timed-out)")
                if len(xg) > 0:
                    xg = [row.copy() for row in xg]; ug = [row.copy() for row in ug]
                    self.tree.add_node(self.node_seq[-1], self.goal, None, xg, ug)
                    self.node_seq.append(self.tree.size - 1)
                    self.x_seq.extend(xg)
                    self.u_seq.extend(ug)
                    self.t_seq = np.arange(len(self.x_seq)) * self.dt
            if self.printing:
                print("Tree size: {0}\nETA: {1} s".format(self.tree.size, np.round(self.T, 2)))
            self._prepare_interpolators()

        elif outcome == 'timeout':                                  # planner.py:311-328
            colmask = np.isinf(self.constraints.goal_buffer)
            self.node_seq, self.x_seq, self.u_seq = self.tree.plan(None, self.xguide, colmask)
            self.T = len(self.x_seq) * self.dt
            self.t_seq = np.arange(len(self.x_seq)) * self.dt
            if self.printing:
                print("Didn't reach goal.\nTree size: {0}\nETA: {1} s".format(self.tree.size, np.round(self.T, 2)))
            self._prepare_interpolators()

        self.stats['gpu_launches'] = eng.launch_count() - launches0
        self.stats['host_closeout_s'] = time.perf_counter() - t_loop1
        self.stats['h2d_bytes'] = eng.h2d_bytes - h2d0
        self.stats['d2h_bytes'] = eng.d2h_bytes - d2h0
        if outcome == 'killed' or self.tree.size > self.max_nodes:  # planner.py:330-336
            if self.printing:
                print("Plan update terminated abruptly!")
            self.killed = False
            return False
        return True

    def _get_exchange(self, eng):
        if self._exchange is None:
            import torch
            _, nbytes = eng.wave_records()
            dev = torch.device('cuda', eng.device)
            if self.exchange_mode == 'peer':
                self._exchange = _waves.PeerExchange(self.comm, nbytes, dev, eng)
            else:
                eng.set_peers(0, 0, [], 0)
                self._exchange = _waves.Exchange(self.comm, nbytes, dev)
        return self._exchange

    def _send_view(self, eng):
        """Zero-copy uint8 view of the engine's record buffer (fixed address for the engine's lifetime)."""
        if getattr(self, '_send_cache', None) is None or self._send_cache[0] is not eng:
            import torch
            from ._native import DevicePointer
            ptr, nbytes = eng.wave_records()
            self._send_cache = (eng, torch.as_tensor(DevicePointer(ptr, nbytes), device=torch.device('cuda', eng.device)))
        return self._send_cache[1]

    # ---------------------------------------------------------- interpolators
    def _prepare_interpolators(self):
        """planner.py:451-464."""
        if len(self.x_seq) == 1:
            self.get_state = lambda t: self.x_seq[0]
            self.get_effort = lambda t: np.zeros(self.ncontrols)
        else:
            self.get_state = scipy.interpolate.interp1d(
                self.t_seq, np.array(self.x_seq), axis=0, assume_sorted=True,
                bounds_error=False, fill_value=self.x_seq[-1][:])
            self.get_effort = scipy.interpolate.interp1d(
                self.t_seq, np.array(self.u_seq), axis=0, assume_sorted=True,
                bounds_error=False, fill_value=self.u_seq[-1][:])

    # ---------------------------------------------------------------- setters
    def set_goal(self, goal):
        """Modifies the goal state and region (planner.py:468-487)."""
        if goal is None:
            self.goal = None
            return
        if len(goal) != self.nstates:
            raise ValueError("The goal state must have same dimensionality as state space.")
        self.goal = np.array(goal, dtype=np.float64)
        self.goal_region = [(self.goal[i] - b, self.goal[i] + b) for i, b in enumerate(self.constraints.goal_buffer)]
        self.plan_reached_goal = False

    def set_runtime(self, min_time=None, max_time=None, max_nodes=None, sys_time=None):
        """Arguments not given are not modified (planner.py:491-513)."""
        if min_time is not None:
            self.min_time = min_time
        if max_time is not None:
            self.max_time = max_time
        if self.min_time > self.max_time:
            raise ValueError("The min_time must be less than or equal to the max_time.")
        if max_nodes is not None:
            self.max_nodes = max_nodes
        if sys_time is not None:
            if not hasattr(sys_time, '__call__'):
                raise ValueError("Expected sys_time to be a function.")
            self.sys_time = sys_time

    def set_resolution(self, horizon=None, dt=None, FPR=None, error_tol=None):
        """Arguments not given are not modified (planner.py:517-553)."""
        if horizon is not None:
            self.horizon = horizon
        if dt is not None:
            self.dt = dt
        if FPR is not None:
            self.FPR = FPR
        if error_tol is not None:
            if np.shape(error_tol) in [(), (self.nstates,)]:
                self.error_tol = np.abs(error_tol).astype(np.float64)
            else:
                raise ValueError("Shape of error_tol must be scalar or length of state.")
        if hasattr(self.horizon, '__contains__'):
            if len(self.horizon) != 2:
                raise ValueError("Expected horizon to be tuple (min, max) or a single scalar.")
            if self.horizon[0] < self.dt:
                raise ValueError("The minimum horizon must be at least as big as dt.")
            if self.horizon[0] >= self.horizon[1]:
                raise ValueError("A horizon range tuple must be given as (min, max) where min < max.")
            self.horizon_iters = 1
            self.hspan = np.divide(self.horizon, self.dt).astype(np.int64)
            self.hfactor = int(2)
        elif self.horizon >= self.dt:
            self.horizon_iters = int(self.horizon / self.dt)
            self.hspan = (self.horizon_iters, self.horizon_iters)
            self.hfactor = 0
        else:
            raise ValueError("The horizon must be at least as big as dt.")

    def set_system(self, dynamics=None, lqr=None, constraints=None, erf=None):
        """Arguments not given are not modified; dynamics and lqr change together (planner.py:557-592)."""
        if dynamics is not None or lqr is not None:
            if not hasattr(dynamics, '__call__'):
                raise ValueError("Expected dynamics to be a function.")
            if not hasattr(lqr, '__call__'):
                raise ValueError("Expected lqr to be a function.")
            self.dynamics = dynamics
            self.lqr = lqr
        if constraints is not None:
            if not isinstance(constraints, Constraints):
                raise ValueError("Expected constraints to be an instance of the Constraints class.")
            self.constraints = constraints
            self.nstates = self.constraints.nstates
            self.ncontrols = self.constraints.ncontrols
        if erf is not None:
            if not hasattr(erf, '__call__'):
                raise ValueError("Expected erf to be a function.")
            self.erf = erf
        self._system = _models.bundle_of(self.dynamics, self.lqr, getattr(self, 'erf', np.subtract),
                                         self.constraints.is_feasible, self.nstates, self.ncontrols)
        if self._system.nstates != self.nstates or self._system.ncontrols != self.ncontrols:
            raise ValueError("constraints say nstates={}, ncontrols={} but the dynamics model has {} and {}".format(
                self.nstates, self.ncontrols, self._system.nstates, self._system.ncontrols))
        self.plan_reached_goal = False

    # -------------------------------------------------------------- kill flag
    def kill_update(self):
        """Abruptly terminates update_plan at its next wave boundary (planner.py:596-601); callable from
        any thread."""
        self.killed = True
        eng = self._engine
        if eng is not None:
            try:
                eng.request_stop(_waves.STOP_KILL)      # the device-scheduled loop polls this word
            except Exception:
                pass

    def unkill(self):
        """planner.py:605-610."""
        self.killed = False

    def visualize(self, dx, dy):
        """planner.py:614-624."""
        if hasattr(self, 'node_seq'):
            self.tree.visualize(dx, dy, node_seq=self.node_seq)
        else:
            print("There is no plan to visualize!")
