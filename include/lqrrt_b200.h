/*
 * lqrrt_b200 -- C ABI of the B200 tree-expansion engine.
 *
 * This is the drop-in boundary for the hot path of jnez71/lqRRT: the body of
 * the `while True:` loop of Planner.update_plan (reference
 * lqrrt/planner.py:233-290) and the functions it calls.  The reference is pure
 * Python and has no FFI of its own; each entry point below names the reference
 * code it replaces, and INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - every function returns int: 0 = LQ_OK, < 0 = error; nothing throws.
 *     lq_last_error(h) returns a NUL-terminated message owned by the handle
 *     (or a static string when h is NULL).
 *   - `h_*` pointers are HOST memory valid for the duration of the call and are
 *     always float64 / int32 regardless of the engine dtype (the reference's
 *     own types).  `d_*` pointers are DEVICE memory owned by the caller (e.g. a
 *     torch tensor's data_ptr) holding the engine dtype (float or double).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - matrices are row-major; node-major host arrays are (count x n).
 *   - one engine = one GPU = one host thread at a time (the reference runs
 *     update_plan on the caller's thread, planner.py:104).
 */
#ifndef LQRRT_B200_H
#define LQRRT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQ_ABI_VERSION 2
#define LQ_MAX_N 12          /* states   */
#define LQ_MAX_M 6           /* controls */
#define LQ_MAX_PRIMS 8

enum { LQ_OK = 0, LQ_ERR_ARG = -1, LQ_ERR_CUDA = -2, LQ_ERR_STATE = -3, LQ_ERR_UNSUPPORTED = -4 };
enum { LQ_F32 = 0, LQ_F64 = 1 };

/* dynamics(x, u, dt) -- reference contract planner.py:35-37 */
enum { LQ_SYS_BOAT = 1,              /* demos/demo_boat_{novice,intermediate,advanced}.py dynamics */
       LQ_SYS_CAR = 2,               /* demos/demo_car.py:46-72 */
       LQ_SYS_PENDULUM2 = 3,         /* synthetic config C1 */
       LQ_SYS_DOUBLE_INTEGRATOR = 4  /* synthetic config C5 */ };

/* (S, K) = lqr(x, u) -- reference contract planner.py:39-42 */
enum { LQ_LQR_PD_ROTATED = 1,        /* demo_boat_novice.py:86-103, demo_car.py:95-113 */
       LQ_LQR_CONSTANT = 2,          /* demo_pendulum.py:119-126 */
       LQ_LQR_FD_RICCATI = 3         /* north_star's built-in provider (not in the reference): central-difference Jacobians of
                                        the device dynamics, `steps` backward Riccati iterations from P = Q.
                                        lqr_params = Q (n x n), R (m x m), eps, steps[, snap]: snap > 0 rounds every Jacobian
                                        entry to a multiple of snap before the sweep (0 / absent: as differenced) */ };

/* is_feasible(x, u) primitives -- reference contract constraints.py:29,53-61 */
enum { LQ_FEAS_STATE_BOX = 1,        /* demo_boat_advanced.py:211-213 */
       LQ_FEAS_EFFORT_BOX = 2,       /* demo_pendulum.py:154-157 */
       LQ_FEAS_STATE_REGION = 3,     /* config C1 obstacle region */
       LQ_FEAS_CIRCLES_POINT = 4,    /* demo_boat_novice.py:160-164 */
       LQ_FEAS_CIRCLES_CLOUD = 5,    /* demo_boat_intermediate.py:198-210, demo_car.py:168-180 */
       LQ_FEAS_BOXES_CLOUD = 6,      /* config C2 boxes */
       LQ_FEAS_SPHERES_POINT = 7,    /* config C5 spheres */
       LQ_FEAS_OGRID_CLOUD = 8       /* demos/lqrrt_ros/nodes/lqrrt_node.py:719-745: hull vertices vs an occupancy grid;
                                        obs = grid rows x cols (row-major, n_obs = rows), lo = {cols, cells per metre,
                                        origin x, origin y, threshold} */ };

typedef struct lq_feas_prim {
    int32_t kind;
    int32_t n_obs;                 /* rows of obs */
    int32_t n_verts;               /* columns of verts (2 x V) */
    int32_t flags;                 /* bit0: demo_car.py:175 doubled-centre vertex */
    int32_t dims[4];               /* state indices: x, y, (z | heading) */
    double inflate;
    const double* obs;             /* n_obs x {3: cx cy r | 4: cx cy hx hy | 4: cx cy cz r} */
    const double* verts;           /* 2 x n_verts, row-major */
    const double* lo;              /* box primitives: n (state) or m (effort) entries */
    const double* hi;
} lq_feas_prim;

typedef struct lq_config {
    int32_t abi_version;           /* LQ_ABI_VERSION */
    int32_t device;                /* CUDA device ordinal */
    int32_t dtype;                 /* LQ_F32 | LQ_F64: arithmetic type of every kernel */
    int32_t n, m;                  /* nstates, ncontrols (constraints.py:32-33) */
    int32_t sys_kind, lqr_kind;
    int32_t n_sys_params, n_lqr_params;
    const double* sys_params;      /* blob layouts: lqrrt_b200/models.py */
    const double* lqr_params;
    uint32_t angle_mask;           /* erf: bit i set => dim i wraps (demo_boat_novice.py:106-118) */
    int32_t n_prims;
    const lq_feas_prim* prims;
    int32_t horizon_iters;         /* planner.py:549; with an adaptive horizon: hspan_hi, the longest edge */
    int32_t hfactor;               /* planner.py:538-551: 0 = fixed horizon, 2 = adaptive-horizon heuristic */
    int32_t hspan_lo, hspan_hi;    /* planner.py:546: clip range of horizon_iters in steps */
    double dt;                     /* planner.py:527 */
    double FPR;                    /* planner.py:530; int(FPR*len) is evaluated in float64 on the device */
    double error_tol[LQ_MAX_N];    /* planner.py:532-536 (inf allowed) */
    int32_t node_capacity;         /* tree rows held in HBM */
    int32_t wave_capacity;         /* most samples one lq_wave_expand may take on this GPU */
    int32_t tries_capacity;        /* most sampler tries (planner.py:188-191) */
} lq_config;

typedef struct lq_wave_stats {
    int32_t tree_size;             /* after commit */
    int32_t n_sampled;             /* samples expanded by this wave (all ranks) */
    int32_t n_accepted;            /* nodes appended */
    int32_t n_goal;                /* goal-region nodes so far this plan */
    int32_t best_node;             /* node of the fastest plan so far, -1 if none (planner.py:276-281) */
    int32_t best_cost;             /* its cumulative steps; T = (1 + best_cost) * dt (planner.py:273) */
    int32_t full;                  /* 1 if node_capacity truncated this wave */
    int32_t tries_used;            /* sum over samples of sampler tries */
    int32_t flags;                 /* `flags` rank 0 passed to its lq_wave_expand (host stop protocol) */
    int32_t horizon_iters;         /* planner-level horizon after this wave's votes (adaptive horizon, planner.py:421-424) */
    int32_t reserved[2];
} lq_wave_stats;

typedef struct lq_engine lq_engine;

/* ---- user-compiled dynamics ----------------------
 * The reference takes `dynamics` as a Python closure (planner.py:35-37); a kernel cannot call Python, so a
 * vehicle that is not one of the compiled models above is supplied as CUDA source for one explicit step
 * (lqrrt_b200/plugin.py UserDynamics).  That source is compiled by nvcc against this engine's own headers
 * into a small shared library whose factory builds the same lq_engine (same kernels, same C ABI) around
 * the user's model; the library registers it here and lq_create dispatches any sys_kind >=
 * LQ_SYS_PLUGIN_BASE to the factory registered for (sys_kind, dtype, lqr_kind).  The factory returns NULL
 * and fills err / *rc on failure; it must not throw. */
#define LQ_SYS_PLUGIN_BASE 1000
typedef lq_engine* (*lq_plugin_factory)(const lq_config* cfg, char* err, int err_len, int* rc);
int lq_register_plugin(int sys_kind, int dtype, int lqr_kind, lq_plugin_factory make);

/* ---- lifetime ----------------------------------------------------------- */
int lq_abi_version(void);
int lq_create(const lq_config* cfg, lq_engine** out);       /* replaces Planner.__init__/set_system/set_resolution, planner.py:85-100,517-592 */
int lq_destroy(lq_engine* h);
const char* lq_last_error(const lq_engine* h);

/* ---- problem state ------------------------------------------------------ */
/* Planner.set_goal, planner.py:468-487: goal and goal region (goal +- buffer, open box). */
int lq_set_goal(lq_engine* h, const double* h_goal, const double* h_buffer);
/* Adaptive horizon only (planner.py:418-425): the planner-level horizon_iters the next wave starts from.
 * Every rollout of a wave starts from this value; its doublings/halving are applied in sample order at
 * commit (wave = 1 is the reference's sequential behaviour) and reported back in lq_wave_stats. */
int lq_set_horizon(lq_engine* h, int horizon_iters, void* stream);
/* default sampler setup, planner.py:178-198: box, goal bias, tries limit. */
int lq_set_sampling(lq_engine* h, const double* h_lo, const double* h_hi, const double* h_goal_bias, int tries);
/* Tree.__init__, tree.py:50-73 + planner.py:172-173: new tree seeded at x0, K0 = lqr(x0, 0); clears ignores/best plan. */
int lq_reset(lq_engine* h, const double* h_x0, void* stream);
/* test/bench hook: replace the tree with `count` given nodes (K_i = lqr(x_i, 0)). */
int lq_load_nodes(lq_engine* h, int count, const double* h_state, const int32_t* h_parent,
                  const int32_t* h_cost, void* stream);
int lq_tree_size(lq_engine* h, int* out);                   /* Tree.size, tree.py:73 (synchronises the stream) */

/* ---- the hot path, stage by stage ---------------------------------------- */
/* _costs_to_go + argsort/argmin selection, planner.py:239-247,340-350.
 * d_queries: n x W (dim-major) engine dtype.  d_idx: W int32, d_cost: W engine dtype (may be NULL). */
int lq_nearest(lq_engine* h, const void* d_queries, int W, int use_ignore,
               int32_t* d_idx, void* d_cost, void* stream);
/* _steer(ID, xtar, force_arrive=False), planner.py:354-438, for W (node, target) pairs.
 * d_targets: n x W dim-major.  Outputs (may be NULL): d_len W int32;
 * d_xseq W x H x n, d_useq W x H x m (sample-major) engine dtype. */
int lq_steer(lq_engine* h, const int32_t* d_ids, const void* d_targets, int W,
             void* d_xseq, void* d_useq, int32_t* d_len, void* stream);

/* One wave of the expand loop, planner.py:233-283, for samples [first, first+count) of a global
 * wave of `wave_size` samples: sample (planner.py:201-211; Philox stream keyed by seed/wave_ctr/sample,
 * or h_samples: count x n float64 when non-NULL, or h_candidates below) -> nearest -> steer, whose
 * epilogue writes every sample's result (edge length; parent, state and gain if a node came out) into the
 * engine's exchange segment -- or straight into every peer's, see lq_wave_set_peers.  Nothing is appended yet. */
int lq_wave_expand(lq_engine* h, uint64_t seed, uint32_t wave_ctr, int first, int count,
                   const double* h_samples, int use_ignore, int flags, void* stream);
/* Reference-stream sampling (planner.py:201-211 on NumPy's MT19937): h_candidates holds
 * `n_candidates` consecutive tries (n_candidates x n float64) in stream order; the engine tests
 * feasibility of all of them, then walks them in order giving each of `count` samples up to `tries`
 * tries.  *h_consumed receives how many candidates were used (the host advances its RNG by that). */
int lq_wave_expand_candidates(lq_engine* h, int count, int tries, const double* h_candidates,
                              int n_candidates, int use_ignore, int* h_consumed, void* stream);
/* `flags` above travels in the record header; after lq_wave_commit every rank reads segment 0's
 * value back in lq_wave_stats.flags -- the host uses it to agree on when to stop without an extra
 * collective (the reference's stop rules are wall-clock based, planner.py:286-311). */
/* The exchange segment of the last lq_wave_expand: device pointer + fixed byte size (the allgather payload):
 * a small header, one meta word per sample of the slice and one record row (state, gain K) + parent per
 * sample that produced a node -- uncompacted, lq_wave_commit derives the node ids.  Edge trajectories stay on
 * the GPU that rolled them out and are read back from their owner (lq_download_edges). */
int lq_wave_records(lq_engine* h, void** d_records, uint64_t* bytes);
/* Tree.add_node for every accepted sample of `n_segments` exchange segments laid end to end
 * (segment r = rank r's lq_wave_records), in segment order then sample order = global sample order;
 * then goal test / ignores / best plan (planner.py:253-283, tree.py:77-96).  d_records may be NULL for
 * the engine's own buffer (single GPU).  `my_segment` is the index of this engine's own segment (its
 * rank): the edges of those nodes are stored here, every node remembers its owner segment. */
int lq_wave_commit(lq_engine* h, const void* d_records, int n_segments, int my_segment, void* stream);
/* Fused rollout + exchange over NVLink.  d_peer_bufs[p] is rank p's receive buffer mapped into this process
 * (peer-mapped symmetric memory, e.g. torch.distributed._symmetric_memory), all of `bytes` >=
 * 2*world*segment + 2*world*4 bytes and zero-initialised.  After this call the rollout kernel stores its
 * results directly into every peer's buffer (16-byte stores) and its last CTA raises an arrival flag, and
 * lq_wave_commit(h, NULL, world, rank, stream) waits for all flags of the wave inside the kernel -- no
 * collective call per wave.  world <= 1 turns it off. */
int lq_wave_set_peers(lq_engine* h, int world, int rank, void* const* d_peer_bufs, uint64_t bytes);
/* Blocks until the last commit finished and returns its counters. */
int lq_wave_stats_get(lq_engine* h, lq_wave_stats* out);

/* ---- the whole expand loop of one plan, scheduled on the device ------------------------------------------
 * Planner.update_plan's `while True:` (planner.py:233-311) as one call: waves of Philox samples are sized on the
 * device (W = clamp(int(growth * N), wave_min, wave_max), or wave_fixed), enqueued `depth` waves ahead of the
 * counters the host has seen, and the scan of the part of the tree a wave cannot have changed runs on a side
 * stream under the previous wave's rollouts and commit (lag 1) -- or whole waves overlap pairwise (lag 2, see
 * `lag`).  Wave k is sized from the tree size wave k-1 started with (waves 0 and 1 from the initial tree); the
 * nodes it appends are exactly those of the stepwise calls lq_wave_expand / lq_wave_commit with the same sizes
 * (and the same lag).
 *   on_wave   called on the caller's thread once per committed wave, in order, with that wave's counters;
 *             returns 0 to go on, 1 to stop (planner.py:293/311: done) or 2 (planner.py:289: killed).  The
 *             request reaches the device through mapped host memory; `flags` of the counters tell when it
 *             took effect (a few waves later) -- on several GPUs rank 0's request is the one that counts and
 *             every rank stops after the same wave.  May be NULL (runs until the tree is full).
 *   world, rank   > 1 GPUs: lq_wave_set_peers must have been called; every rank calls lq_plan_run with
 *             the same arguments. */
typedef int (*lq_wave_callback)(void* user, const lq_wave_stats* wave);
typedef struct lq_plan_args {
    uint64_t seed;                 /* Philox key */
    uint32_t wave_ctr0;            /* Philox wave counter of the first wave (keeps counting across plans) */
    int32_t wave_fixed;            /* > 0: every wave has this many samples */
    int32_t wave_min, wave_max;    /* adaptive wave size bounds */
    double growth;
    int32_t use_ignore;            /* pruning (planner.py:239-245) */
    int32_t world, rank;
    int32_t depth;                 /* waves in flight ahead of the host (0 = default) */
    int32_t lag;                   /* 1: wave k searches the tree as committed by wave k-1 (speculative old-tree scan + delta scan);
                                      2: pipelined waves -- wave k searches the tree (nodes and ignore marks) as committed by wave
                                      k-2, so the rollouts of two consecutive waves are in flight together.  0 = 1. */
    lq_wave_callback on_wave;
    void* user;
} lq_plan_args;
typedef struct lq_plan_summary {
    int32_t waves;                 /* waves whose counters were delivered (incl. the no-op tail after a stop) */
    int32_t waves_enqueued;
    int32_t stop_code;             /* 0: ran out of capacity only; else the stop request / full flag seen by the device */
    int32_t reserved;
    int64_t samples;
    double device_ms;              /* CUDA-event time of the wave pipeline on its own stream */
    double host_enqueue_ms;        /* host time spent enqueuing waves / waiting for counters / in on_wave: if the first */
    double host_wait_ms;           /* approaches device_ms the pipeline is fed too slowly (launch-bound) */
    double host_callback_ms;
    lq_wave_stats last;
} lq_plan_summary;
int lq_plan_run(lq_engine* h, const lq_plan_args* args, lq_plan_summary* out, void* stream);
/* The wave-sizing rule of the device schedule, as plain host functions (what the commit kernel evaluates):
 * W = wave_fixed if > 0, else clamp(int(growth * tree_size), wave_min, wave_max) in float64; and the contiguous
 * slice [first, first + count) of a wave of W samples that `rank` of `world` expands (rank-order concatenation of
 * the slices is sample order).  Mirrors lqrrt_b200/waves.py wave_size / partition. */
int lq_policy_wave(int wave_fixed, int wave_min, int wave_max, double growth, int tree_size);
int lq_policy_slice(int W, int rank, int world, int32_t* first, int32_t* count);
/* The host word the rollout kernels poll for a stop request (mapped pinned memory): another thread may set it to
 * 1 (done) or 2 (kill) at any time, e.g. Planner.kill_update (planner.py:596-601); lq_plan_run clears it on entry. */
int lq_stop_word(lq_engine* h, int32_t** out);

/* closest-to-guide, planner.py:313-318: argmin_i d_i' Sg d_i, Sg = lqr(xguide,0).S with the
 * columns of dims where h_colmask[i] != 0 zeroed. */
int lq_closest(lq_engine* h, const double* h_xguide, const uint8_t* h_colmask, int* h_id, void* stream);
/* _steer(ID, xtar, force_arrive=True), planner.py:296,399-410, with the wall-clock limit replaced by
 * max_steps; h_x: max_steps x n, h_u: max_steps x m. */
int lq_steer_force(lq_engine* h, int node, const double* h_xtar, int max_steps,
                   double* h_x, double* h_u, int* h_len, void* stream);
/* Tree.add_node with explicit content (finish_on_goal, planner.py:299). */
int lq_append_node(lq_engine* h, int parent, const double* h_state, const double* h_x, const double* h_u,
                   int len, void* stream);
/* planner.py:270 for externally chosen ids (tests). */
int lq_mark_ignore(lq_engine* h, const int32_t* h_ids, int count, void* stream);

/* Tree.climb, tree.py:100-117: node ids from the seed (first) to `node` (last), walked on the device so a
 * plan read-out does not have to download the whole parent array. */
int lq_climb(lq_engine* h, int node, int32_t* h_ids, int max_len, int* h_len, void* stream);

/* The plan in one call (planner.py:266-281 when the goal was reached, :313-324 otherwise): node >= 0 climbs from that
 * node; node < 0 first finds the node closest to h_xguide as lq_closest does.  h_ids receives the path seed-first
 * (*h_len entries, at most max_len); h_x (max_len x H x n), h_u (max_len x H x m), h_edge_len, when non-NULL, the
 * edges of those nodes in the same order (zeros beyond an edge's length and for edges held by another GPU, as
 * lq_download_edges).  Two host synchronisations in total. */
int lq_plan_readout(lq_engine* h, int node, const double* h_xguide, const uint8_t* h_colmask, int max_len,
                    int32_t* h_ids, int* h_len, double* h_x, double* h_u, int32_t* h_edge_len, void* stream);

/* ---- read-back (Tree.state / pID / lqr / x_seq / u_seq, tree.py:62-70) --- */
int lq_download_nodes(lq_engine* h, int first, int count, double* h_state, int32_t* h_parent,
                      int32_t* h_cost, int32_t* h_edge_len, uint8_t* h_ignore, uint8_t* h_owner, double* h_K, void* stream);
/* edges of `count` nodes listed in h_ids: h_x count x H x n, h_u count x H x m.  Rows past the edge length are
 * zero, and so is the whole edge of a node owned by another GPU (sum the per-rank answers to assemble a plan).
 * h_len (may be NULL): the recorded edge lengths (known on every rank). */
int lq_download_edges(lq_engine* h, const int32_t* h_ids, int count, double* h_x, double* h_u, int32_t* h_len, void* stream);

/* ---- the user callbacks themselves, batched (planner.py:35-59) ------------ */
int lq_eval_dynamics(lq_engine* h, int count, const double* h_x, const double* h_u, double dt, double* h_xnext);
int lq_eval_lqr(lq_engine* h, int count, const double* h_x, const double* h_u, double* h_S, double* h_K);
int lq_eval_erf(lq_engine* h, int count, const double* h_xgoal, const double* h_x, double* h_e);
int lq_eval_feasible(lq_engine* h, int count, const double* h_x, const double* h_u, uint8_t* h_ok);

/* number of kernels this engine has launched since creation (bench.py's gpu_launches) */
int lq_launch_count(lq_engine* h, uint64_t* out);
/* tuning knobs: steer_lanes = threads sharing one rollout (1, 8 or 32; 0 keeps the heuristic),
 * nearest_q = nearest queries per thread (0 = heuristic, 1 = small, >1 = large) */
int lq_set_tuning(lq_engine* h, int steer_lanes, int nearest_q);
/* Engine options.  Every setting produces the same tree, bit for bit; they select how the work is issued.
 *   LQ_OPT_NEAREST_KERNEL  float32 diagonal-metric scan (planner.py:340-350): LQ_NEAREST_SCALAR = one FADD + one
 *                          FFMA per (query, node, dim) and a per-pair compare/select; LQ_NEAREST_PACKED = the same
 *                          arithmetic on the FP32x2 pipe (two queries per instruction) with a two-level argmin.
 *                          float64 engines and dense metrics ignore it.
 *   LQ_OPT_SAMPLE_PREFETCH 1: the device sampler (planner.py:201-211) of a wave is issued on a side stream into
 *                          the second sample buffer, so it overlaps the previous wave's scan and rollouts. */
enum { LQ_OPT_NEAREST_KERNEL = 1, LQ_OPT_SAMPLE_PREFETCH = 2,
       LQ_OPT_NEAREST_CTAS_PER_SM = 3,  /* grid target of the scan in CTAs per SM; 0 = auto (16 for n <= 6, else 8) */
       LQ_OPT_SPECULATIVE_SCAN = 4,     /* lq_plan_run: 1 (default) = old-tree scan of wave k+1 on a side stream under wave k */
       LQ_OPT_FORCE_RESCAN = 5,         /* tests: every query takes the rollout kernel's own rescan path */
       LQ_OPT_STAGE_TABLES = 6,         /* 1 (default) = obstacle tables staged in shared memory by the rollout kernels */
       LQ_OPT_STEER_BLOCK = 7,          /* CTA size of the wave rollout kernel: 0 = auto, 32, 128 */
       LQ_OPT_PDL = 8,                  /* 1 (default) = rollout and commit kernels of a wave are launched programmatically
                                           dependent: they start (and stage their tables) while the kernel in front drains */
       /* Indexed nearest search (csrc/grid.cuh): the tree keeps a uniform grid over its first two state dims and a
        * query only evaluates the nodes of cells whose lower cost bound is below its best cost so far -- the same
        * pair arithmetic and tie rule as the brute-force scan, hence the same answer bit for bit.  Needs a constant
        * diagonal metric with positive weights on two un-wrapped leading dims; 1 = use it for scans of at least
        * LQ_OPT_GRID_MIN_KPAIRS x 1024 (query, node) pairs, 0 (default) = always scan everything.  Measured on C4
        * (profiles/README.md): wave-grown trees are so clustered in (x, y) -- thousands of nodes per cell that differ
        * in heading and velocity only -- that a query still evaluates ~30 000 candidates, one dependent load at a
        * time, and loses to the brute-force scan; kept for sparse, wide trees. */
       LQ_OPT_NEAREST_GRID = 9,
       LQ_OPT_GRID_MIN_KPAIRS = 10,
       LQ_OPT_SAMPLE_AHEAD = 11,        /* pipelined waves: 1 (default) = once the wave size is known to the host (top of the
                                           policy, fixed waves) the sampler runs on a side stream ahead of the wave's chain */
       /* State-dependent lqr providers (FD-Jacobian + Riccati) whose metric S = lqr(x, 0).S (planner.py:348) does not in
        * fact depend on x: lq_set_sampling probes S at 12 points of the sample box; if all agree with their mean within
        * the finite-difference noise of the dtype (3e-2 relative in float32, 1e-7 in float64) and no state dim wraps, that
        * mean is THE metric: factored S = U^T U on the host, the tree keeps z = U v, and the search is the diagonal scan
        * over z (cost = |U q - U v|^2) instead of one Riccati sweep per sample plus a dense scan.  1 (default) = do so;
        * 0 = always evaluate S per sample.  Unlike the other options this one changes float32 results within that noise
        * (it removes it); lq_get_option reports whether it is in effect. */
       LQ_OPT_WHITEN_METRIC = 12 };
enum { LQ_NEAREST_SCALAR = 0, LQ_NEAREST_PACKED = 1,
       /* tuning variants, same results: 8 queries per thread (best on the 12-state scan), and the two-level
        * argmin with one-lane-at-a-time arithmetic */
       LQ_NEAREST_PACKED_Q8 = 2, LQ_NEAREST_TWO_LEVEL = 3,
       /* experimental (not measured yet): partial-distance filter -- the cost of the leading position dims is a
        * floating-point lower bound of the full cost, pairs that cannot win skip the other dims; same results */
       LQ_NEAREST_PREFIX = 4 };
int lq_set_option(lq_engine* h, int option, int value);
int lq_get_option(lq_engine* h, int option, int* value);
/* Per-stage CUDA-event timing of the wave chain on the launching stream (bench.py's roofline leg).
 * ms[5]: sample, nearest (old-tree scan, side stream), nearest (delta / whole-tree scan), steer, commit (accumulated);
 * work[4]: nearest pairs sum(W*N),
 * nearest algorithmic bytes sum(W*N*n*sizeof(dtype)), samples, waves. */
int lq_set_profiling(lq_engine* h, int on);
int lq_profile_read(lq_engine* h, double* ms5, double* work4, int reset);

#ifdef __cplusplus
}
#endif
#endif
