#!/usr/bin/env python
"""
bench.py -- tree expansions/s (nodes added per second) of the expand loop, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c4]

Workload (config.workload): BASELINE.json configs[3], "demo_boat_advanced.py 6-DoF boat, 500
obstacles, 100k-node tree"  -- the configuration the metric's 1/2/4/8-GPU
scaling is quoted on and the north_star's ">= 1e6 expansions/s on the 6-DoF boat" target.  One STEP =
one whole plan: a new tree grown from the seed until it holds max_nodes nodes, through the public
`Planner.update_plan` call (sample -> nearest -> steer -> commit waves).  Every step uses a new Philox
seed (different samples, same obstacle field), so nothing is cached between steps; the tree (2.4 MB of
states) is L2-resident by design, which `config.l2` states.

  value   nodes added / device time of the wave loop (CUDA events on the launching stream; samples are
          generated on the device, so every input is HBM-resident), max over ranks
  e2e     the same plans timed on the host clock around the whole update_plan call: host-side problem
          inputs (seed state, goal, sample box) copied in, per-wave counters read back, and the plan
          (node chain + edges) downloaded and interpolators built -- bytes counted from those copies
  roofline  the nearest-neighbour scan (K1): algorithmic bytes sum_k W_k * N_k * n * 4
          over its CUDA-event time inside the same timed steps, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the sequential NumPy oracle port of the reference loop, one core, a bounded wall budget
          on the same obstacle field (the unmodified reference is Python and cannot travel to the box;
          the port is pinned bit-for-bit to it by tests/test_oracle_*.py)

--impl reference times that CPU port (the reference's algorithm on the host cores: one independent
replica per core, different seeds, nodes/s summed) and prints the same JSON shape.
"""
import argparse
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
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    ncu --set full capture (profiles/ncu_nearest_*.txt); None if there is none."""
    import glob
    import re
    best = None
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
            best = float(np.mean(vals))
    return best


def make_problem(name):
    from lqrrt_b200 import configs
    return configs.BY_NAME[name](0)


def workload_name(prob, name):
    return {
        'c4': "C4 demo_boat_advanced 6-DoF boat, 500 circular obstacles, 200 m arena, tree grown to 100k nodes",
        'c3': "C3 demo_boat_novice 6-DoF boat, 50 circular obstacles, tree grown to max_nodes",
        'c2': "C2 demo_car 5-state car, 20 box obstacles, tree grown to max_nodes",
        'c5': "C5 12-state double integrator, 5000 spheres, FD-Riccati lqr, tree grown to 1M nodes",
        'c5k': "C5 12-state double integrator, 5000 spheres, constant LQR gain, tree grown to 1M nodes",
    }.get(name, prob.name)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for k, nm in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[2 + k].lower().startswith('active') for r in self.rows):
                reasons.append(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------
# CPU side: the sequential oracle port
# ---------------------------------------------------------------------------------------------
def cpu_growth(name, seed, budget_s):
    """Sequential reference-algorithm growth from an empty tree for a wall budget; returns
    (nodes added, iterations, seconds)."""
    from oracle import models as om
    from oracle.seqplanner import SeqPlanner
    prob = make_problem(name)
    dyn, lqr, erf, feas = om.make_callbacks(prob.system, prob.dt)
    sp = SeqPlanner(dyn, lqr, prob.nstates, prob.ncontrols, prob.goal_buffer, feas, prob.horizon, prob.dt,
                    prob.FPR, prob.error_tol, erf, min_time=budget_s, max_time=budget_s, max_nodes=prob.max_nodes,
                    goal0=prob.goal)
    np.random.seed(seed)
    t0 = time.perf_counter()
    sp.update_plan(prob.x0, prob.sample_space, goal_bias=prob.goal_bias, xrand_gen=10)
    dt = time.perf_counter() - t0
    return sp.tree.size - 1, sp.iterations, dt


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
    """The reference's own CPU algorithm on the host cores (oracle port; the Python reference cannot be
    shipped to the GPU box).  One independent replica per usable core, nodes/s summed."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    prob = make_problem(args.config)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    cores = max(1, min(cores, 64))
    budget = float(args.cpu_seconds)
    steps = max(1, args.steps)
    per = []
    ctx = mp.get_context('fork')
    t_all = time.perf_counter()
    with ctx.Pool(cores) as pool:
        for s in range(args.warmup + steps):
            if s < args.warmup:
                res = pool.map(_cpu_worker, [(args.config, 1000 + c, min(1.0, budget)) for c in range(cores)])
                continue
            res = pool.map(_cpu_worker, [(args.config, 100 * s + c, budget) for c in range(cores)])
            per.append((sum(r[0] / r[2] for r in res), sum(r[0] for r in res), max(r[2] for r in res)))
    value = float(np.mean([p[0] for p in per]))
    ms = float(np.mean([p[2] for p in per])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(prob, args.config), "note": "sequential expand loop (one sample per iteration), "
                   "trees stay at N~1e2-1e3 within the budget; the reference cannot reach 100k nodes (BASELINE.md section 2)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "{} independent replicas x {:.0f} s of sequential growth from an empty tree, summed".format(cores, budget)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
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
                                nearest_kernel=args.nearest, prefetch=args.prefetch)
    packed = planner.nearest_kernel != 'scalar' and args.dtype == 'f32'
    # Issue slots and FMA-pipe cycles per (query, node) pair of the scan's inner loop, n dims of which w wrap
    # (cuobjdump -sass of csrc/nearest.cuh, checked against ncu's instruction counts in profiles/README.md):
    #   scalar: FADD+FFMA per dim, 4 per wrapped dim, compare + 2 selects, LDS/loop 0.5      -> 17.5 at n=6, w=1
    #   packed: FADD2+FFMA2 per dim per TWO pairs, wrapped dims scalar, FMNMX + 3/8 per pair,
    #           LDS.128 n/8 per pair                                                          -> 11.3 at n=6, w=1
    #   FMA pipe: FFMA2/FADD2 hold it 2 cycles, so both forms cost 2(n-w) + 3w cycles per pair -> 13
    n_dims = prob.nstates
    n_wrap = bin(prob.system.erf.mask(n_dims)).count('1')
    if packed:
        slots_per_pair = (n_dims - n_wrap) + 4 * n_wrap + 1.4 + (n_dims + 1) // 2 * 2 / 8.0 + 0.2
    else:
        slots_per_pair = 2 * (n_dims - n_wrap) + 4 * n_wrap + 3.5
    pipe_cycles_per_pair = 2 * (n_dims - n_wrap) + 3 * n_wrap

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
        phases = np.zeros(3); reached = 0; depth = []; plan_T = []
        for s in range(args.steps):
            host_s, st = one_plan(s)
            host_times.append(host_s); dev_ms.append(st['device_ms']); nodes.append(st['nodes'] - 1)
            samples.append(st['samples']); waves.append(st['waves']); launches += st['gpu_launches']
            d2h += st['d2h_bytes']; h2d += st['h2d_bytes']
            phases += [st['host_setup_s'], st['host_loop_s'], st['host_closeout_s']]
            reached += int(planner.plan_reached_goal); depth.append(len(planner.node_seq)); plan_T.append(float(planner.T))
        barrier()
    prof = planner._engine.profile_read()

    tot_dev = float(np.sum(dev_ms)) / 1e3
    tot_host = float(np.sum(host_times))
    if world > 1:
        t = torch.tensor([tot_dev, tot_host], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_dev, tot_host = float(t[0]), float(t[1])
    total_nodes = float(np.sum(nodes))
    value = total_nodes / tot_dev
    e2e = total_nodes / tot_host

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm = float(peaks.get('hbm_gbs', 6650.0))
        near_s = prof['ms']['nearest'] / 1e3
        elt = 4 if args.dtype == 'f32' else 8
        achieved = prof['nearest_bytes'] / near_s / 1e9 if near_s > 0 else 0.0
        stage_total = sum(prof['ms'].values())
        # secondary bound: FP32 issue rate.  ~16 issue slots per (query, node) pair for n=6 (csrc/nearest.cuh)
        pairs_per_s = prof['pairs'] / near_s if near_s > 0 else 0.0
        cpu = None
        if not args.no_cpu:
            n_add, iters, secs = cpu_growth(args.config, 0, float(args.cpu_seconds))
            cpu = {"value": n_add / secs, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "{:.0f} s of sequential growth from an empty tree on the same obstacle field "
                             "({} iterations, {} nodes)".format(secs, iters, n_add)}
            if args.cpu_frozen:
                tr = planner.tree
                cpu["frozen_tree"] = cpu_frozen_tree(args.config, tr.state, tr.pID)
                cpu["frozen_tree_note"] = "one loop iteration of the CPU port on the first N nodes of the GPU-grown tree " \
                                          "(BASELINE.md protocol B); a GPU wave at that size expands thousands of samples"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_dev / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(prob, args.config), "max_nodes": int(prob.max_nodes),
                       "wave": "adaptive clamp({:g}*N, 32, {})".format(args.growth, wave_max), "wave_per_gpu_max": wave_max // world,
                       "waves_per_step": float(np.mean(waves)), "samples_per_step": float(np.mean(samples)),
                       "yield": total_nodes / float(np.sum(samples)), "sampler": "philox4x32-10 on device, new seed per step",
                       "nearest_kernel": planner.nearest_kernel, "sample_prefetch": planner.prefetch_active,
                       "plans_reaching_goal": "{}/{}".format(reached, args.steps), "plan_depth_edges": float(np.mean(depth)),
                       "plan_duration_s": float(np.mean(plan_T)),
                       "note": "tree depth grows by at most one edge per wave, so the throughput-maximal wave cap "
                               "fills 100k nodes before the 283 m-away goal is reached; --wave 512 reaches it "
                               "(profiles/README.md has both)",
                       "l2": "tree states (<= 2.4 MB) and obstacle tables are L2-resident by design; every step "
                             "grows a new tree from new samples (no reuse between timed steps)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": tot_host / args.steps * 1e3,
                    "host_ms_per_step": dict(zip(("setup", "wave_loop", "closeout"), (phases / args.steps * 1e3).round(3).tolist()))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "nearest_diag_pk_kernel" if packed else "nearest_diag_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm if hbm else None, "traffic": ncu_traffic(),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes": prof['nearest_bytes'], "kernel_seconds": near_s,
                         "share_of_device_time": (prof['ms']['nearest'] / stage_total) if stage_total else None,
                         "pairs_per_s": pairs_per_s,
                         "fp32_issue_frac": pairs_per_s * slots_per_pair / (148 * 128 * 1.965e9),
                         "issue_slots_per_pair": slots_per_pair,
                         "fma_pipe_frac": pairs_per_s * pipe_cycles_per_pair / (148 * 128 * 1.965e9),
                         "fma_pipe_cycles_per_pair": pipe_cycles_per_pair,
                         "note": "algorithmic bytes = sum_k W_k*N_k*n*sizeof(dtype) (SURVEY.md 8d); node tiles are reused "
                                 "by every query of a CTA from shared memory and the tree is L2-resident, so frac > 1 is "
                                 "expected and the binding limit is the FP32 pipe: fp32_issue_frac = pairs_per_s x issue_slots_per_pair "
                                 "/ (148 SMs x 128 lanes x 1.965 GHz) bounds the scalar kernel (ncu: 90 % of issue slots), "
                                 "fma_pipe_frac = pairs_per_s x fma_pipe_cycles_per_pair / the same denominator bounds the packed "
                                 "kernel (ncu: FMA pipe 67.6 % of cycles, math-pipe-throttle the top stall); traffic = measured DRAM bytes "
                                 "per launch from the committed ncu capture"},
            "stage_ms": prof['ms'],
            "clocks": clk.summary(),
        }
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
    ap.add_argument('--wave', type=int, default=8192, help="largest wave per GPU")
    ap.add_argument('--scaling', default='strong', choices=['weak', 'strong'],
                    help="strong: same global wave schedule for every N (bit-identical tree); weak: 8192 samples per GPU per wave")
    ap.add_argument('--growth', type=float, default=2.0,
                    help="adaptive wave size W = clamp(growth * tree size, 32, --wave); 1.0 is the Planner's own default "
                         "(deeper tree per sample, ~40 more small waves per 100k-node plan)")
    ap.add_argument('--nearest', default=None, help="nearest-scan kernel (default: the planner's)")
    ap.add_argument('--prefetch', default=None, type=int, choices=[0, 1], help="side-stream sampler prefetch (default: the planner's)")
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
