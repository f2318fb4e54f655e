#!/usr/bin/env python
"""
bench.py -- tree expansions/s (nodes added per second) of the expand loop, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c4]

Workload (config.workload): BASELINE.json configs[3], "demo_boat_advanced.py 6-DoF boat, 500
obstacles, 100k-node tree"  -- the configuration the metric's 1/2/4/8-GPU
scaling is quoted on and the north_star's ">= 1e6 expansions/s on the 6-DoF boat" target.  One STEP =
one whole plan: a new tree grown from the seed until it holds max_nodes nodes, through the public
`Planner.update_plan` call (the device-scheduled wave pipeline, lq_plan_run).  Every step uses a new
Philox seed (different samples, same obstacle field), so nothing is cached between steps; the tree
(2.4 MB of states) is L2-resident by design, which `config.l2` states.

  value   nodes added / device time of the wave pipeline (CUDA events on the pipeline's own chain stream;
          samples are generated on the device, so every input is HBM-resident), max over ranks
  e2e     the same plans timed on the host clock around the whole update_plan call: host-side problem
          inputs (seed state, goal, sample box) copied in, per-wave counters read back, and the plan
          (node chain + edges) downloaded and interpolators built -- bytes counted from those copies
  roofline  the nearest-neighbour scan (K1).  It is bound by the FP32 pipe, not by HBM (the tree is L2-resident
          and every staged tile serves 512 queries): `achieved` = algorithmic flops (3n per (query, node) pair,
          SURVEY.md 8d secondary bound) over the scan kernels' CUDA-event time, `peak` = SMs x 128 lanes x 2 x the
          SM clock sampled during the run; the HBM view the contract also asks for (algorithmic bytes
          sum_k W_k N_k n 4 over the same time, against MEASURED_PEAKS.json) is kept as `hbm_effective_*`
  plan    a second leg outside the timed region: the same field planned with the goal-reaching wave policy
          (small waves: tree depth grows one edge per wave) until the first plan to the goal 283 m away
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, copied from /root/reference by build()) on one host core
          for a bounded wall budget on the same obstacle field; falls back to the oracle port (bit-pinned to the
          reference by tests/test_oracle_*.py) where oracle/_ref did not travel, `kind` says which

--impl reference times that CPU implementation on all host cores (one independent replica per core, different
seeds, nodes/s summed; every step a bounded sample so that the whole run stays within a few minutes) and prints the
same JSON shape.  `config` is the same object in both arms (the workload and how the GPU arm expands it); what a run
measured besides the metric (waves, samples, yield, plan statistics) is under `run`.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "tree_expansions_per_s"
UNIT = "nodes/s"


def ncu_traffic():
    """(bytes, file): dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, parsed from
    the newest committed ncu --set full capture (profiles/ncu_nearest_*.txt) -- a constant of that capture, not
    something this run measured; (None, None) if there is none."""
    import glob
    import re
    best = (None, None)
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'ncu_nearest_*.txt'))):
        vals = []
        rd = wr = None
        for line in open(path):
            m = re.match(r'dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)', line)
            if m:
                v = float(m.group(2)) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(m.group(3), 1)
                if m.group(1) == 'read':
                    rd = v
                else:
                    wr = v
                if rd is not None and wr is not None:
                    vals.append(rd + wr); rd = wr = None
        if vals:
            best = (float(np.mean(vals)), os.path.relpath(path, ROOT))
    return best


def make_problem(name):
    from lqrrt_b200 import configs
    return configs.BY_NAME[name](0)


def workload_name(prob, name):
    return {
        'c4': "C4 demo_boat_advanced 6-DoF boat, 500 circular obstacles, 200 m arena, tree grown to 100k nodes",
        'c3': "C3 demo_boat_novice 6-DoF boat, 50 circular obstacles, tree grown to max_nodes",
        'c2': "C2 demo_car 5-state car, 20 box obstacles, tree grown to max_nodes",
        'c1': "C1 2-state pendulum swing-up, FD-Riccati lqr, 1 forbidden region, tree grown to max_nodes",
        'c5': "C5 12-state double integrator, 5000 spheres, FD-Jacobian + Riccati lqr (Jacobians snapped to 0.025), tree grown to 1M nodes",
        'c5k': "C5 12-state double integrator, 5000 spheres, constant LQR gain, tree grown to 1M nodes",
    }.get(name, prob.name)


def static_config(args, prob, world):
    """What both arms of the benchmark print as `config`: the workload and how the GPU arm expands it -- nothing measured
    (per-run figures go under `run`), so the two lines carry the same object."""
    wave_max = args.wave * world if args.scaling == 'weak' else args.wave
    lag = args.lag
    nearest = args.nearest or os.environ.get('LQRRT_NEAREST') or 'packed'
    return {"workload": workload_name(prob, args.config), "max_nodes": int(prob.max_nodes),
            "wave": "adaptive clamp({:g}*N, 32, {}), wave k sized from the tree wave k-1 started with".format(args.growth, wave_max),
            "wave_per_gpu_max": wave_max // world, "schedule": args.schedule or 'device',
            "lag": "{} ({})".format(lag, "pipelined waves: wave k searches the tree committed by wave k-2" if lag == 2
                                    else "wave k searches the tree committed by wave k-1"),
            "sampler": "philox4x32-10 on device, new seed per step", "nearest_kernel": nearest,
            "note": "throughput workload: the tree is filled to max_nodes with the largest waves; tree depth grows by "
                    "at most one edge per wave, so it fills before the 283 m-away goal is reached -- the `plan` "
                    "object times the goal-reaching policy on the same field.  The reference arm (--impl reference) runs "
                    "the reference's sequential loop (one sample per iteration) on the same obstacle field for a bounded "
                    "wall budget; it stays at N ~ 1e2-1e3 nodes and cannot reach max_nodes (BASELINE.md section 2)",
            "l2": "tree states (<= 2.4 MB) and obstacle tables are L2-resident by design; every step "
                  "grows a new tree from new samples (no reuse between timed steps)"}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process every 2 ms (a 100k-node plan
    takes ~7 ms, nvidia-smi takes ~40 ms per call), nvidia-smi as the fallback where NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []            # (sm MHz, max MHz, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap)
        self.source = None
        self.stop = threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            ids = [v.strip() for v in vis.split(',') if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def open_nvml(self):
        """NVML set-up (tens of ms) happens before the timed region, not inside the sampling thread."""
        import pynvml as nv
        nv.nvmlInit()
        self.nv = nv
        self.h = nv.nvmlDeviceGetHandleByIndex(self.physical_index())
        self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        self.reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.sample_nvml()
        self.rows = []

    def sample_nvml(self):
        nv = self.nv
        r = int(self.reasons(self.h))
        self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), self.mx,
                          bool(r & nv.nvmlClocksThrottleReasonHwSlowdown), bool(r & nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                          bool(r & nv.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & nv.nvmlClocksThrottleReasonSwPowerCap)))

    def run_nvml(self):
        self.source = 'nvml'
        while not self.stop.is_set():
            self.sample_nvml()
            self.stop.wait(0.002)

    def run_smi(self):
        self.source = 'nvidia-smi'
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.physical_index()), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 6:
                    self.rows.append((float(parts[0]), float(parts[1])) + tuple(p.lower().startswith('active') for p in parts[2:6]))
            except Exception:
                pass
            self.stop.wait(0.05)

    def run(self):
        try:
            if self.nv is None:
                raise RuntimeError("no NVML")
            self.run_nvml()
        except Exception:
            if not self.stop.is_set():
                self.run_smi()

    def __enter__(self):
        self.nv = None
        try:
            self.open_nvml()
        except Exception:
            self.nv = None
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(r[0] for r in self.rows)
        reasons = [nm for k, nm in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"))
                   if any(r[2 + k] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


# ---------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref) or, where it did not travel, the oracle port
# ---------------------------------------------------------------------------------------------
REF_DEMO = {'c4': 'demo_boat_advanced', 'c3': 'demo_boat_novice'}


def reference_available(name):
    from oracle import build_ref
    return name in REF_DEMO and build_ref.available()


def cpu_growth(name, seed, budget_s):
    """Sequential growth from an empty tree for a wall budget on the config's obstacle field; returns
    (nodes added, iterations or None, seconds, kind, reached, T)."""
    prob = make_problem(name)
    if reference_available(name):
        # the reference's own Planner and the demo's own dynamics / lqr / erf / is_feasible; only the obstacle
        # table, goal and sample box are the benchmark's (the demo file's module globals, set from the outside)
        from oracle import build_ref
        ref = build_ref.load()
        ns = build_ref.demo_model(REF_DEMO[name], 0)
        ns['obs'] = np.array(prob.obs, dtype=np.float64)
        cons = ref.Constraints(nstates=prob.nstates, ncontrols=prob.ncontrols, goal_buffer=prob.goal_buffer,
                               is_feasible=ns['is_feasible'])
        planner = ref.Planner(ns['dynamics'], ns['lqr'], cons, horizon=prob.horizon, dt=prob.dt, FPR=prob.FPR,
                              error_tol=prob.error_tol, erf=ns['erf'], min_time=budget_s, max_time=budget_s,
                              max_nodes=prob.max_nodes, goal0=prob.goal, printing=False)
        np.random.seed(seed)
        t0 = time.perf_counter()
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
        dt = time.perf_counter() - t0
        return planner.tree.size - 1, None, dt, 'reference', bool(planner.plan_reached_goal), float(planner.T)
    from oracle import models as om
    from oracle.seqplanner import SeqPlanner
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    sp = SeqPlanner(dyn, lqr, prob.nstates, prob.ncontrols, prob.goal_buffer, feas, prob.horizon, prob.dt,
                    prob.FPR, prob.error_tol, erf, min_time=budget_s, max_time=budget_s, max_nodes=prob.max_nodes,
                    goal0=prob.goal)
    np.random.seed(seed)
    t0 = time.perf_counter()
    sp.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    dt = time.perf_counter() - t0
    return sp.tree.size - 1, sp.iterations, dt, 'port', bool(sp.plan_reached_goal), float(sp.T)


def cpu_frozen_tree(name, state, parent, sizes=((1000, 20), (10000, 8), (100000, 3))):
    """BASELINE.md protocol B: inject the first N nodes of a GPU-grown tree into the sequential CPU port and
    time whole loop iterations (sample -> _costs_to_go -> selection -> _steer) at that fixed tree size."""
    from oracle import models as om
    from oracle.seqplanner import SeqPlanner, SeqTree, costs_to_go, default_sampler
    prob = make_problem(name)
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    sp = SeqPlanner(dyn, lqr, prob.nstates, prob.ncontrols, prob.goal_buffer, feas, prob.horizon, prob.dt,
                    prob.FPR, prob.error_tol, erf, goal0=prob.goal)
    np.random.seed(0)
    gen = default_sampler(prob.sample_space, prob.goal_bias, sp.goal, feas, prob.ncontrols, 10)
    out = []
    zero = np.zeros(prob.ncontrols)
    for N, iters in sizes:
        if N > len(state):
            continue
        tree = SeqTree(state[0], lqr(state[0], zero))
        tree.state = np.array(state[:N], dtype=np.float64)
        tree.pID = list(parent[:N]); tree.size = N
        tree.lqr = [lqr(x, zero) for x in tree.state]
        sp.tree = tree
        t0 = time.perf_counter()
        for _ in range(iters):
            xr = gen(None)
            c = costs_to_go(sp.lqr, sp.erf_v, sp.ncontrols, tree.state, xr)
            sp.steer(int(np.argsort(c)[0]), xr)
        dt = (time.perf_counter() - t0) / iters
        out.append({"tree_size": N, "ms_per_iteration": dt * 1e3, "iterations_per_s": 1.0 / dt})
    return out


def _cpu_worker(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    return cpu_growth(*args)


def run_reference_arm(args):
    """The reference's own CPU implementation on the host cores: one independent replica per usable core (the
    reference is single-threaded), nodes/s summed."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    prob = make_problem(args.config)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    cores = max(1, min(cores, 64))
    steps = max(1, args.steps)
    # every step is a bounded sample of the workload; the whole run stays within a few minutes whatever --steps is
    budget = min(float(args.cpu_seconds), max(2.0, 120.0 / steps))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    per = []
    kind = 'port'
    ctx = mp.get_context('fork')
    t_all = time.perf_counter()
    with ctx.Pool(cores) as pool:
        for s in range(args.warmup + steps):
            if s < args.warmup:
                res = pool.map(_cpu_worker, [(args.config, 1000 + c, min(1.0, budget)) for c in range(cores)])
                continue
            res = pool.map(_cpu_worker, [(args.config, 100 * s + c, budget) for c in range(cores)])
            kind = res[0][3]
            per.append((sum(r[0] / r[2] for r in res), sum(r[0] for r in res), max(r[2] for r in res)))
    value = float(np.mean([p[0] for p in per]))
    ms = float(np.mean([p[2] for p in per])) * 1e3
    what = "unmodified reference (oracle/_ref: lqrrt.Planner + the demo's own model functions)" if kind == 'reference' \
        else "oracle port of the reference loop (oracle/_ref absent on this box)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": static_config(args, prob, world),
        "run": {"implementation": what, "seconds_per_replica_per_step": budget,
                "note": "sequential expand loop (one sample per iteration); trees stay at N~1e2-1e3 within the budget"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "{} independent replicas x {:.0f} s of sequential growth from an empty tree, summed".format(cores, budget)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def tree_digest(tree):
    """sha1 over node states (float64 as downloaded), parents and cumulative costs: equal digests = the same tree,
    node for node and bit for bit.  The wave schedule does not depend on the number of GPUs, so the driver's
    1/2/4/8-GPU runs must all print the same digest."""
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(tree.state).tobytes())
    h.update(np.asarray(tree.pID, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(tree.cost).astype(np.int32).tobytes())
    return h.hexdigest()[:16]


def scan_alone(planner, prob, launches=20):
    """The nearest-neighbour scan by itself: saturated searches (one full wave of fresh queries) of the tree the last timed
    step grew, back to back on the launching stream with nothing else on the GPU, CUDA events around them.  Inside the
    pipeline two waves' scans share the machine with each other and with rollouts and commits, so the per-stage event
    times there are not kernel durations."""
    import torch
    eng = planner._engine
    n = eng.n
    size = int(planner.tree.size)
    W = int(planner._wave_capacity())
    ss = np.array(prob.sample_space, dtype=np.float64)
    Q = ss[:, 0] + (ss[:, 1] - ss[:, 0]) * np.random.RandomState(0).rand(W, n)
    dt = torch.float32 if eng.dtype == 'f32' else torch.float64
    q = torch.tensor(np.ascontiguousarray(Q.T), dtype=dt, device='cuda')
    idx = torch.empty(W, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.nearest(q.data_ptr(), W, False, idx.data_ptr(), None, st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        eng.nearest(q.data_ptr(), W, False, idx.data_ptr(), None, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / launches
    return {"queries": W, "tree_size": size, "launches": launches, "ms_per_launch": ms,
            "pairs_per_s": W * float(size) / (ms * 1e-3)}


def goal_leg(prob, args, comm, local, plans):
    """Plans that actually reach the goal: the same obstacle field with the goal-reaching wave policy (waves of at most
    `--plan-wave` samples -- depth grows one edge per wave and the goal is >= 130 edges away), stopping at the first
    plan found (min_time = 0).  Outside the timed region of `value`."""
    planner = prob.make_planner(dtype=args.dtype, wave=None, wave_max=args.plan_wave, wave_growth=1.0, max_nodes=prob.max_nodes,
                                min_time=0.0, max_time=10.0, comm=comm, device=local, lag=1,
                                exchange=os.environ.get('LQRRT_EXCHANGE', 'peer'))
    ms = []; T = []; nodes = []; waves = []; reached = 0
    for s in range(plans + 1):
        planner.seed = 5000 + s
        t0 = time.perf_counter()
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
        dt = time.perf_counter() - t0
        if s == 0:
            continue                       # warm-up (engine construction)
        if planner.plan_reached_goal:
            reached += 1
            ms.append(dt * 1e3); T.append(float(planner.T)); nodes.append(planner.tree.size); waves.append(planner.stats['waves'])
    out = {"policy": "adaptive clamp(N, 32, {}) waves, lag 1, stop at the first plan (min_time 0)".format(args.plan_wave),
           "plans_reaching_goal": "{}/{}".format(reached, plans)}
    if reached:
        out.update({"time_to_first_plan_ms": float(np.mean(ms)), "time_to_first_plan_ms_max": float(np.max(ms)),
                    "plan_duration_s": float(np.mean(T)), "tree_size_at_plan": float(np.mean(nodes)),
                    "waves_to_plan": float(np.mean(waves))})
    if planner._engine is not None:
        planner._engine.close()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm = True
    if args.gpus != world and rank == 0:
        sys.stderr.write("note: --gpus {} but WORLD_SIZE {}; using WORLD_SIZE\n".format(args.gpus, world))

    prob = make_problem(args.config)
    wave_max = args.wave * world if args.scaling == 'weak' else args.wave
    planner = prob.make_planner(dtype=args.dtype, wave=None, wave_max=wave_max, wave_growth=args.growth, max_nodes=prob.max_nodes,
                                min_time=1e9, max_time=1e9, comm=comm, device=local,
                                exchange=os.environ.get('LQRRT_EXCHANGE', 'peer'),
                                nearest_kernel=args.nearest, prefetch=args.prefetch, schedule=args.schedule, lag=args.lag)
    packed = planner.nearest_kernel != 'scalar' and args.dtype == 'f32'
    # FMA-pipe cycles per (query, node) pair of the scan's inner loop, n dims of which w wrap (cuobjdump -sass of
    # csrc/nearest.cuh, checked against ncu's instruction counts in profiles/README.md): FFMA2/FADD2 hold the pipe 2
    # cycles, so the packed and the scalar form both cost 2(n-w) + 3w cycles per pair -> 13 at n = 6, w = 1
    n_dims = prob.nstates
    n_wrap = bin(prob.system.erf.mask(n_dims)).count('1')
    pipe_cycles_per_pair = 2 * (n_dims - n_wrap) + 3 * n_wrap
    flops_per_pair = 3 * n_dims                       # n sub, n mul, n add for a diagonal metric

    def one_plan(seed):
        planner.seed = seed
        planner.profile = True
        t0 = time.perf_counter()
        planner.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias)
        host_s = time.perf_counter() - t0
        return host_s, dict(planner.stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        one_plan(9000 + w)
    if planner._engine is not None:
        planner._engine.profile_read(reset=True)

    barrier()
    with ClockSampler(local) as clk:
        host_times = []; dev_ms = []; nodes = []; samples = []; waves = []; launches = 0; d2h = 0; h2d = 0
        phases = np.zeros(3); reached = 0; depth = []; plan_T = []; hostpipe = np.zeros(3)
        for s in range(args.steps):
            host_s, st = one_plan(s)
            host_times.append(host_s); dev_ms.append(st['device_ms']); nodes.append(st['nodes'] - 1)
            samples.append(st['samples']); waves.append(st['waves']); launches += st['gpu_launches']
            d2h += st['d2h_bytes']; h2d += st['h2d_bytes']
            phases += [st['host_setup_s'], st['host_loop_s'], st['host_closeout_s']]
            hp = st.get('host_pipeline_ms')
            if hp:
                hostpipe += [hp['enqueue_ms'], hp['wait_ms'], hp['callback_ms']]
            reached += int(planner.plan_reached_goal); depth.append(len(planner.node_seq)); plan_T.append(float(planner.T))
        barrier()
    prof = planner._engine.profile_read()
    digest = tree_digest(planner.tree)
    alone = scan_alone(planner, prob) if not hasattr(prob.system.lqr, 'steps') else None     # (state-dependent metrics: in-pipeline figure only)
    schedule_used = planner.stats.get('schedule')

    tot_dev = float(np.sum(dev_ms)) / 1e3
    tot_host = float(np.sum(host_times))
    if world > 1:
        t = torch.tensor([tot_dev, tot_host], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_dev, tot_host = float(t[0]), float(t[1])
        # every rank must hold the same tree
        dg = [None] * world
        dist.all_gather_object(dg, digest)
        if len(set(dg)) != 1:
            raise SystemExit("ranks disagree on the tree: {}".format(dg))
    total_nodes = float(np.sum(nodes))
    value = total_nodes / tot_dev
    e2e = total_nodes / tot_host

    plan = None
    if args.plan_leg and args.config in ('c4', 'c3', 'demo_boat_novice'):
        plan = goal_leg(prob, args, comm, local, args.plan_leg)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm = float(peaks.get('hbm_gbs', 6650.0))
        clocks = clk.summary()
        sm_hz = (clocks.get('sm_mhz') or float(peaks.get('sm_max_mhz', 1965.0))) * 1e6
        fp32_peak = 148 * 128 * 2 * sm_hz / 1e12                 # TFLOP/s of the FP32 pipes at the clock of this run
        near_s = (prof['ms']['nearest_old'] + prof['ms']['nearest_delta']) / 1e3
        pairs_per_s = prof['pairs'] / near_s if near_s > 0 else 0.0
        achieved_pipe = pairs_per_s * flops_per_pair / 1e12
        # primary figure: the scan alone on the GPU (scan_alone); the in-pipeline figure (overlapped stage times) beside it
        achieved = alone['pairs_per_s'] * flops_per_pair / 1e12 if alone else achieved_pipe
        hbm_eff = prof['nearest_bytes'] / near_s / 1e9 if near_s > 0 else 0.0
        stage_total = sum(prof['ms'].values())
        traffic, traffic_file = ncu_traffic()
        cpu = None
        if not args.no_cpu:
            n_add, iters, secs, kind, cpu_reached, cpu_T = cpu_growth(args.config, 0, float(args.cpu_seconds))
            cpu = {"value": n_add / secs, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": "{:.0f} s of sequential growth from an empty tree on the same obstacle field ({} nodes{})".format(
                       secs, n_add, "" if iters is None else ", {} iterations".format(iters)),
                   "reached_goal": cpu_reached, "plan_duration_s": cpu_T if np.isfinite(cpu_T) else None}
            if args.cpu_frozen:
                tr = planner.tree
                cpu["frozen_tree"] = cpu_frozen_tree(args.config, tr.state, tr.pID)
                cpu["frozen_tree_note"] = "one loop iteration of the CPU port on the first N nodes of the GPU-grown tree " \
                                          "(BASELINE.md protocol B); a GPU wave at that size expands thousands of samples"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_dev / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "tree_digest": digest,
            "config": static_config(args, prob, world),
            "run": {"schedule": schedule_used, "lag": planner.stats.get('lag'), "nearest_kernel": planner.nearest_kernel,
                    "waves_per_step": float(np.mean(waves)), "samples_per_step": float(np.mean(samples)),
                    "yield": total_nodes / float(np.sum(samples)),
                    "plans_reaching_goal": "{}/{}".format(reached, args.steps), "plan_depth_edges": float(np.mean(depth)),
                    "plan_duration_s": float(np.mean(plan_T))},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": tot_host / args.steps * 1e3,
                    "host_ms_per_step": dict(zip(("setup", "wave_loop", "closeout"), (phases / args.steps * 1e3).round(3).tolist())),
                    "wave_loop_host_ms_per_step": dict(zip(("enqueue", "wait_for_counters", "callback"), (hostpipe / args.steps).round(3).tolist()))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32_pipe", "kernel": "nearest_diag_pk_kernel" if packed else "nearest_diag_kernel",
                         "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
                         "peak_source": "148 SMs x 128 FP32 lanes x 2 x SM clock sampled during this run ({:.0f} MHz)".format(sm_hz / 1e6),
                         "measured_on": ("{launches} saturated launches ({queries} queries x {tree_size} nodes) after the timed steps, nothing "
                                         "else on the GPU, CUDA events on the launching stream".format(**alone)) if alone else
                                        "per-stage CUDA events inside the pipeline",
                         "alone": alone,
                         "flops_per_pair": flops_per_pair, "pairs_per_s": alone['pairs_per_s'] if alone else pairs_per_s,
                         "fma_pipe_frac": (alone['pairs_per_s'] if alone else pairs_per_s) * pipe_cycles_per_pair / (148 * 128 * sm_hz),
                         "in_pipeline": {"achieved": achieved_pipe, "frac": achieved_pipe / fp32_peak if fp32_peak else None,
                                         "pairs_per_s": pairs_per_s, "kernel_seconds": near_s,
                                         "fma_pipe_frac": pairs_per_s * pipe_cycles_per_pair / (148 * 128 * sm_hz),
                                         "note": "sum of event-bracketed scan stages over the timed steps; with pipelined waves two scans "
                                                 "(and the other wave's rollouts / commit) share the GPU, so these seconds overlap each other"},
                         "share_of_device_time": (near_s * 1e3 / stage_total) if stage_total else None,
                         "fma_pipe_cycles_per_pair": pipe_cycles_per_pair,
                         "hbm_effective_gbs": hbm_eff, "hbm_peak_gbs": hbm, "hbm_effective_frac": hbm_eff / hbm if hbm else None,
                         "hbm_peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes": prof['nearest_bytes'],
                         "traffic": traffic, "traffic_source": "constant from {} (ncu --set full of one launch), not measured by this run".format(traffic_file) if traffic_file else None,
                         "note": "the scan's tree is L2-resident and every staged node tile serves 512 queries, so HBM does not bound it "
                                 "(hbm_effective_frac = algorithmic bytes sum_k W_k*N_k*n*sizeof(dtype) over the kernel time against the HBM peak is > 1 "
                                 "by construction; measured DRAM traffic is `traffic` per launch).  The binding unit is the FP32 pipe: frac = "
                                 "3n flops per (query, node) pair over the pipe peak; fma_pipe_frac counts the cycles the packed FADD2/FFMA2 "
                                 "instructions hold the pipe (ncu: 67.6 % of cycles on a saturated launch).  Kernel time = CUDA events around "
                                 "the scan launches on the streams they run on; the old-tree scan shares the GPU with the previous wave's "
                                 "rollouts and commit, which is the point of it."},
            "stage_ms": prof['ms'],
            "stage_note": "per-stage CUDA-event sums over the timed steps; nearest_old and sample run on the scan stream concurrently with "
                          "the previous wave's nearest_delta -> steer -> commit chain, so the stages do not add up to the step time",
            "clocks": clocks,
        }
        if plan is not None:
            line["plan"] = plan
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c4')
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
    ap.add_argument('--wave', type=int, default=16384, help="largest wave per GPU")
    ap.add_argument('--scaling', default='strong', choices=['weak', 'strong'],
                    help="strong: same global wave schedule for every N (bit-identical tree); weak: 8192 samples per GPU per wave")
    ap.add_argument('--growth', type=float, default=2.0,
                    help="adaptive wave size W = clamp(growth * tree size, 32, --wave); 1.0 is the Planner's own default "
                         "(deeper tree per sample, more small waves per 100k-node plan)")
    ap.add_argument('--nearest', default=None, help="nearest-scan kernel (default: the planner's)")
    ap.add_argument('--prefetch', default=None, type=int, choices=[0, 1], help="side-stream sampler prefetch of the host-scheduled loop")
    ap.add_argument('--schedule', default=None, choices=['device', 'host'], help="wave scheduling (default: the planner's, 'device')")
    ap.add_argument('--lag', type=int, default=2, choices=[1, 2],
                    help="wave semantics of the throughput leg: 2 = pipelined waves (wave k searches the tree of wave k-2; the "
                         "Planner's own default is 1, which the goal-reaching leg uses)")
    ap.add_argument('--plan-leg', type=int, default=10, help="goal-reaching plans timed after the throughput steps (0 = skip)")
    ap.add_argument('--plan-wave', type=int, default=512, help="wave cap of the goal-reaching policy")
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-frozen', action='store_true', help="also time the CPU port on frozen trees of 1e3/1e4/1e5 nodes")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
