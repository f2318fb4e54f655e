// The engine behind the C ABI (include/lqrrt_b200.h): owns the tree in HBM as
// structure-of-arrays and runs the reference's expand loop
// (lqrrt/planner.py:233-290) one wave at a time:
//   sample (K3) -> nearest (K1) -> steer (K2) -> pack -> [allgather by the host] -> commit (K4)
// Templated on arithmetic type and on the compiled device model.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <chrono>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#include <algorithm>

#include "models.cuh"
#include "nearest.cuh"
#include "steer.cuh"
#include "sample.cuh"
#include "commit.cuh"
#include "grid.cuh"

struct lq_engine {
    std::string err;
    uint64_t launches = 0;
    int device = 0;
    virtual ~lq_engine() {}
    virtual int set_goal(const double* goal, const double* buf) = 0;
    virtual int set_horizon(int horizon_iters, cudaStream_t st) = 0;
    virtual int set_sampling(const double* lo, const double* hi, const double* bias, int tries) = 0;
    virtual int reset(const double* x0, cudaStream_t st) = 0;
    virtual int load_nodes(int count, const double* state, const int* parent, const int* cost, cudaStream_t st) = 0;
    virtual int tree_size(int* out) = 0;
    virtual int nearest(const void* d_q, int W, int use_ignore, int* d_idx, void* d_cost, cudaStream_t st) = 0;
    virtual int steer(const int* d_ids, const void* d_t, int W, void* d_x, void* d_u, int* d_len, cudaStream_t st) = 0;
    virtual int wave_expand(uint64_t seed, uint32_t wave, int first, int count, const double* h_samples,
                            int use_ignore, int flags, cudaStream_t st) = 0;
    virtual int wave_expand_candidates(int count, int tries, const double* h_cand, int n_cand, int use_ignore,
                                       int* h_consumed, cudaStream_t st) = 0;
    virtual int wave_records(void** d, uint64_t* bytes) = 0;
    virtual int wave_commit(const void* d_records, int n_seg, int my_seg, cudaStream_t st) = 0;
    virtual int wave_stats(lq_wave_stats* out) = 0;
    virtual int plan_run(const lq_plan_args* a, lq_plan_summary* out, cudaStream_t st) = 0;
    virtual int stop_word(int** out) = 0;
    virtual int plan_readout(int node, const double* xg, const uint8_t* colmask, int max_len, int* ids, int* len,
                             double* hx, double* hu, int* elen, cudaStream_t st) = 0;
    virtual int closest(const double* xg, const uint8_t* colmask, int* id, cudaStream_t st) = 0;
    virtual int steer_force(int node, const double* xtar, int max_steps, double* hx, double* hu, int* len, cudaStream_t st) = 0;
    virtual int append_node(int parent, const double* state, const double* hx, const double* hu, int len, cudaStream_t st) = 0;
    virtual int mark_ignore(const int* ids, int count, cudaStream_t st) = 0;
    virtual int climb(int node, int* ids, int max_len, int* len, cudaStream_t st) = 0;
    virtual int download_nodes(int first, int count, double* state, int* parent, int* cost, int* elen,
                               uint8_t* ign, uint8_t* owner, double* K, cudaStream_t st) = 0;
    virtual int download_edges(const int* ids, int count, double* hx, double* hu, int* hlen, cudaStream_t st) = 0;
    virtual int eval_dynamics(int count, const double* x, const double* u, double dt, double* xn) = 0;
    virtual int eval_lqr(int count, const double* x, const double* u, double* S, double* K) = 0;
    virtual int eval_erf(int count, const double* xg, const double* x, double* e) = 0;
    virtual int eval_feasible(int count, const double* x, const double* u, uint8_t* ok) = 0;
    virtual int set_tuning(int steer_lanes, int nearest_q) = 0;
    virtual int set_peers(int world, int rank, void* const* bufs, uint64_t bytes) = 0;
    virtual int set_option(int option, int value) = 0;
    virtual int get_option(int option, int* value) = 0;
    virtual int set_profiling(int on) = 0;
    virtual int profile_read(double* ms, double* work, int reset) = 0;
};

namespace lq {

#define LQ_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char b_[512];                                                                      \
            snprintf(b_, sizeof b_, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, #call); \
            this->err = b_;                                                                    \
            return LQ_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define LQ_FAIL(code, ...)                       \
    do {                                         \
        char b_[512];                            \
        snprintf(b_, sizeof b_, __VA_ARGS__);    \
        this->err = b_;                          \
        return code;                             \
    } while (0)

// <<<>>> with the programmatic-dependent-launch attribute (see nearest.cuh pdl_wait): the kernel may start while
// its predecessor in the stream drains.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_maybe_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl,
                                    Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static_assert(sizeof(lq_wave_stats) == 12 * sizeof(int), "device stats block must mirror lq_wave_stats");
// partial-minimum slots per query: the old-tree scan writes [0, MAX_SPLIT_OLD), the delta scan (or the one full
// scan of a stepwise wave) [MAX_SPLIT_OLD, MAX_SLOTS)
constexpr int MAX_SPLIT_OLD = 384;
constexpr int MAX_SPLIT_NEW = 32;
constexpr int MAX_SLOTS = MAX_SPLIT_OLD + MAX_SPLIT_NEW;
constexpr int MAX_SPLIT = MAX_SPLIT_NEW;
constexpr int NEAR_BLOCK = 128;

// ---- small utility kernels --------------------------------------------------
template <typename T, class Sys, class Lqr>
__global__ void init_gain_kernel(Model<T, Sys, Lqr> md, const T* __restrict__ state, T* __restrict__ K, int cap, int count) {
    constexpr int N = Sys::N, M = Sys::M;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    T x[N], u[M], Kl[M * N];
    for (int d = 0; d < N; d++) x[d] = state[(size_t)d * cap + i];
    for (int r = 0; r < M; r++) u[r] = T(0);
    md.lqr.gain(md.sys, md.dt, x, u, Kl);                         // planner.py:172: lqr(x0, zeros)
    for (int k = 0; k < M * N; k++) K[(size_t)k * cap + i] = Kl[k];
}

template <typename T, class Sys, class Lqr>
__global__ void eval_kernel(Model<T, Sys, Lqr> md, int what, int count, const T* __restrict__ a, const T* __restrict__ b,
                            T dt, T* __restrict__ o1, T* __restrict__ o2, uint8_t* __restrict__ ob) {
    constexpr int N = Sys::N, M = Sys::M;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (what == 0) {          // dynamics: a = x (count x N), b = u (count x M)
        T x[N], u[M], xn[N];
        for (int d = 0; d < N; d++) x[d] = a[i * N + d];
        for (int r = 0; r < M; r++) u[r] = b[i * M + r];
        md.sys.step(x, u, dt, xn);
        for (int d = 0; d < N; d++) o1[i * N + d] = xn[d];
    } else if (what == 1) {   // lqr: S (N x N), K (M x N)
        T x[N], u[M], S[N * N], K[M * N];
        for (int d = 0; d < N; d++) x[d] = a[i * N + d];
        for (int r = 0; r < M; r++) u[r] = b[i * M + r];
        md.lqr.metric(md.sys, md.dt, x, u, S);
        md.lqr.gain(md.sys, md.dt, x, u, K);
        for (int k = 0; k < N * N; k++) o1[(size_t)i * N * N + k] = S[k];
        for (int k = 0; k < M * N; k++) o2[(size_t)i * M * N + k] = K[k];
    } else if (what == 2) {   // erf: a = xgoal, b = x
        T xg[N], x[N], e[N];
        for (int d = 0; d < N; d++) { xg[d] = a[i * N + d]; x[d] = b[i * N + d]; }
        erf_eval<T, N>(md.angle_mask, xg, x, e);
        for (int d = 0; d < N; d++) o1[i * N + d] = e[d];
    } else {                  // is_feasible
        T x[N], u[M];
        for (int d = 0; d < N; d++) x[d] = a[i * N + d];
        for (int r = 0; r < M; r++) u[r] = b[i * M + r];
        ob[i] = feasible_eval<T, N, M>(md.feas, x, u) ? 1 : 0;
    }
}

template <typename T>
__global__ void mark_ignore_kernel(TreeView<T> tv, const int* __restrict__ ids, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int node = ids[i];
    if (node < 0 || node >= *tv.size) return;
    const unsigned bit = 1u << (8 * (node & 3));
    for (int c = 0; c < 2; c++) {
        unsigned* word = (unsigned*)(tv.ignore[c] + (node & ~3));
        if (!(atomicOr(word, bit) & bit)) atomicAdd(tv.n_ignored2[c], 1);
    }
}

template <typename T>
__global__ void gather_edges_kernel(TreeView<T> tv, const int* __restrict__ ids, int count, int my_seg, int H, int HN, int HM,
                                    T* __restrict__ ox, T* __restrict__ ou, int* __restrict__ olen) {
    const int r = blockIdx.x;
    if (r >= count) return;
    const int node = ids[r];
    const int L = min(tv.edge_len[node], H);
    if (threadIdx.x == 0) olen[r] = L;
    // rows past the edge length, and edges held by another GPU, read back as zero
    const int mine = my_seg < 0 || tv.owner[node] == my_seg || node == 0 || (tv.owner[node] == 255 && my_seg == 0);
    const int nx = mine ? L * (HN / H) : 0, nu = mine ? L * (HM / H) : 0;
    for (int k = threadIdx.x; k < HN; k += blockDim.x) ox[(size_t)r * HN + k] = k < nx ? tv.edge_x[(size_t)node * HN + k] : T(0);
    for (int k = threadIdx.x; k < HM; k += blockDim.x) ou[(size_t)r * HM + k] = k < nu ? tv.edge_u[(size_t)node * HM + k] : T(0);
}

// Tree.climb (tree.py:100-117) on the device: follow parents from `node` to the seed; ids come out
// leaf-first, the host reverses them.  One thread: a dependent pointer chase of a few hundred hops.
template <typename T>
__global__ void climb_kernel(TreeView<T> tv, int node, int* __restrict__ ids, int max_len, int* __restrict__ len) {
    int n = 0;
    while (node >= 0 && n < max_len) { ids[n++] = node; node = tv.parent[node]; }
    *len = node >= 0 ? -1 : n;          // -1: path longer than the buffer
}

// climb from the node held in *d_node (the closest-to-guide result) -- lets closest and climb share one host sync
template <typename T>
__global__ void climb_from_kernel(TreeView<T> tv, const int* __restrict__ d_node, int* __restrict__ ids, int max_len, int* __restrict__ len) {
    int node = *d_node, n = 0;
    while (node >= 0 && n < max_len) { ids[n++] = node; node = tv.parent[node]; }
    *len = node >= 0 ? -1 : n;
}

template <typename T, int N> struct SeedArgs { T x[N]; };
template <typename T, int N>
__global__ void seed_kernel(TreeView<T> tv, SeedArgs<T, N> s) {
    if (threadIdx.x < N) tv.state[(size_t)threadIdx.x * tv.cap] = s.x[threadIdx.x];
    if (threadIdx.x == 0) { tv.parent[0] = -1; tv.cost[0] = 0; tv.edge_len[0] = 0; }
}

template <typename T>
__global__ void tree_scalars_kernel(TreeView<T> tv, int size, WaveCtl* ctl) {
    *tv.size = size; *tv.n_goal = 0; *tv.best = ~0ull;
    for (int c = 0; c < 2; c++) { *tv.n_ignored2[c] = 0; *tv.goal_count[c] = 0; }
    ctl->stopped = 0; ctl->acc[0] = ctl->acc[1] = ctl->acc[2] = 0; ctl->done_steer[0] = ctl->done_steer[1] = 0; ctl->done_commit = 0;
    for (int k = 0; k < CTL_RING; k++) ctl->ring[k] = WaveSlot{0, 0, 0, size};
}

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

// ---- the engine -----------------------------------------------------------
template <typename T, class Sys, class Lqr>
struct EngineImpl : lq_engine {
    static constexpr int N = Sys::N, M = Sys::M;
    static constexpr int NS = N * (N + 1) / 2;
    using Md = Model<T, Sys, Lqr>;

    Md md;
    SamplerParams<T, N> sp;
    bool have_goal = false, have_sampling = false;
    int cap = 0, Wc = 0, H = 0, tries_cap = 10;
    int size_ub = 0;                 // host upper bound of the device-side tree size
    int size_ub_exact = 0;           // tree size as of the last stats read (wave-start size for profiling)
    int steer_lanes = 1, near_q = 0; // tuning (0 = heuristic)
    int steer_block = 0;             // CTA size of the wave rollout kernel (0 = heuristic)
    int near_kernel = LQ_NEAREST_SCALAR;   // lq_set_option(LQ_OPT_NEAREST_KERNEL)
    int near_ctas_per_sm = 0;              // lq_set_option(LQ_OPT_NEAREST_CTAS_PER_SM): grid target of the scan, 0 = auto
    int force_rescan = 0;                  // lq_set_option(LQ_OPT_FORCE_RESCAN): tests -- every query takes the rescan path
    int spec_scan = 1;                     // lq_set_option(LQ_OPT_SPECULATIVE_SCAN): old-tree scan on the side stream
    int stage_tables_on = 1;               // lq_set_option(LQ_OPT_STAGE_TABLES)
    int pdl_on = 1;                        // lq_set_option(LQ_OPT_PDL): programmatic dependent launch of rollouts and commit
    // lq_set_option(LQ_OPT_SAMPLE_PREFETCH): sampler of the next wave on a side stream (stepwise driver)
    bool prefetch_on = false;
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_sampled[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    bool consumed_valid[2] = {false, false};
    int sample_parity = 0;
    int cur_parity = 0;              // sample / partials / segment parity of the wave being expanded
    bool metric_diag = true;
    // spatial index of the tree for the indexed nearest search (grid.cuh); kept whenever the metric allows it
    DevBuf<int> g_head, g_next;
    DevBuf<unsigned> g_bits;
    bool grid_ok = false;                  // constant diagonal metric with positive weights on the un-wrapped dims 0, 1
    bool grid_dirty = false;               // the extents changed under a non-empty tree: re-link before the next search
    bool grid_have_extents = false;
    int grid_opt = 0;                      // lq_set_option(LQ_OPT_NEAREST_GRID)
    int grid_min_kpairs = 16384;           // lq_set_option(LQ_OPT_GRID_MIN_KPAIRS)
    DiagMetric<T, N> dmet;
    // State-dependent providers whose metric turns out NOT to depend on the state over the sample box (FD-Riccati on
    // linear dynamics: config C5) are searched through a whitened copy of the tree with the diagonal scans
    // (nearest.cuh whiten_apply): S = mean of the probes = U^T U, cost = |U q - U v|^2.
    DevBuf<T> t_z, w_U, w_zq;
    DiagMetric<T, N> dmet_w;               // identity metric on the whitened copy
    bool whiten_on = false, z_dirty = false;
    int whiten_opt = 1;                    // lq_set_option(LQ_OPT_WHITEN_METRIC)
    double whiten_box[2 * N];              // the sample box the probes were drawn from
    bool whiten_probed = false;
    double Sconst[N * N];
    StagePlan stage_plan;

    // tree: dim-major so that a scan reads each dim as one contiguous stream
    DevBuf<T> t_state, t_K, t_ex, t_eu;
    DevBuf<int> t_parent, t_cost, t_elen, t_scal;
    DevBuf<uint8_t> t_ignore, t_owner;
    DevBuf<int> t_goals;
    static constexpr int GOAL_CAP = 1 << 16;
    int ign_copy = 0;                // copy of the ignore marks the last commit wrote (read-back, stand-alone searches)
    int last_my_seg = 0;             // segment index this GPU committed as (its rank)
    // fused exchange over peer-mapped symmetric memory (lq_wave_set_peers); layout of every rank's buffer:
    // [2 parities][world segments][seg_sz] then [2][world] arrival flags
    int peer_world = 0, peer_rank = 0;
    char* peer_buf[MAX_PEERS] = {};
    unsigned peer_seq = 0;           // waves expanded in peer mode (flags carry it; monotonic for the engine's life)
    DevBuf<unsigned long long> t_best;
    TreeView<T> tv;
    // obstacle tables and their broad-phase grids
    std::vector<DevBuf<T>*> feas_bufs;
    std::vector<DevBuf<int>*> grid_bufs;
    int grid_min_obs = 16;
    bool use_hull_grid = true;
    // wave working set (double-buffered by wave parity where two waves overlap)
    DevBuf<T> w_samples, w_pcost, w_Sp, w_xs, w_us, w_cand, w_goal;
    DevBuf<int> w_pidx, w_len, w_tries, w_pick, w_misc, w_nidx;
    DevBuf<T> w_ncost;
    DevBuf<uint8_t> w_candok;
    DevBuf<int> w_hor;
    DevBuf<char> w_seg;                   // [2 parities] this engine's own exchange segment
    DevBuf<WaveCtl> w_ctl;
    size_t seg_sz = 0;
    // Per-wave counters come back through mapped pinned host memory: the publishing CTA of a commit writes the
    // block and then its sequence number, the host polls the number (no copy, no event on the way).
    static constexpr int RING = 8;
    DevStats* h_ring = nullptr;           // pinned + mapped [RING]
    DevStats* d_ring = nullptr;           // the device's view of it
    int* h_stop = nullptr;                // pinned + mapped: the host's stop request, read by the rollout kernel
    int* d_stop = nullptr;
    lq_wave_stats last_stats;             // what lq_wave_stats_get returns when nothing is pending
    int* h_misc = nullptr;                // pinned
    unsigned q_head = 0, q_tail = 0;      // [q_head, q_tail) = committed, unread (stepwise driver)
    unsigned stat_seq = 0;                // sequence number of the last commit launched
    unsigned q_seq[RING] = {};
    int q_add[RING] = {};                 // most nodes that wave can have appended (for the size bound)
    cudaStream_t q_stream[RING] = {};
    cudaStream_t last_stream = nullptr;
    // the device-scheduled driver's own streams: rollout chain (high priority) and scan / sampler (low priority)
    cudaStream_t s_chain = nullptr, s_scan = nullptr, s_chain2 = nullptr, s_samp = nullptr;
    static constexpr int EVR = 8;
    cudaEvent_t ev_commit[EVR] = {}, ev_scan[EVR] = {}, ev_in = nullptr, ev_begin = nullptr, ev_end = nullptr;
    cudaEvent_t ev_steer[EVR] = {}, ev_samp[EVR] = {};     // pipelined waves: sampler ahead of the chain
    int sample_ahead = 1;                                  // lq_set_option(LQ_OPT_SAMPLE_AHEAD)
    int prof_scan_ev[RING] = {};                           // which event marks the start of the scan stage
    // optional per-stage CUDA-event timing (bench.py roofline): sample | nearest (old-tree + delta) | steer | commit
    bool prof_on = false;
    bool prof_open[RING] = {};
    cudaEvent_t pev[RING][8] = {};
    double prof_ms[5] = {0, 0, 0, 0, 0};   // sample, nearest_old, nearest_delta, steer, commit
    double prof_work[4] = {0, 0, 0, 0};    // nearest pairs (W*N), nearest bytes (W*N*n*sizeof T), samples, waves
    int prof_W[RING] = {};

    ~EngineImpl() override {
        cudaSetDevice(device);
        for (auto* b : feas_bufs) delete b;
        for (auto* b : grid_bufs) delete b;
        if (h_ring) cudaFreeHost(h_ring);
        if (h_stop) cudaFreeHost(h_stop);
        if (h_misc) cudaFreeHost(h_misc);
        for (auto& r : pev) for (auto& e : r) if (e) cudaEventDestroy(e);
        for (auto& e : ev_sampled) if (e) cudaEventDestroy(e);
        for (auto& e : ev_consumed) if (e) cudaEventDestroy(e);
        for (auto& e : ev_commit) if (e) cudaEventDestroy(e);
        for (auto& e : ev_scan) if (e) cudaEventDestroy(e);
        for (auto& e : ev_steer) if (e) cudaEventDestroy(e);
        for (auto& e : ev_samp) if (e) cudaEventDestroy(e);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
        if (aux) cudaStreamDestroy(aux);
        if (s_chain) cudaStreamDestroy(s_chain);
        if (s_scan) cudaStreamDestroy(s_scan);
        if (s_chain2) cudaStreamDestroy(s_chain2);
        if (s_samp) cudaStreamDestroy(s_samp);
    }

    template <typename U>
    int upload(DevBuf<U>& dst, const std::vector<U>& src) {
        LQ_CUDA(dst.alloc(src.size()));
        if (!src.empty()) LQ_CUDA(cudaMemcpy(dst.p, src.data(), src.size() * sizeof(U), cudaMemcpyHostToDevice));
        return LQ_OK;
    }

    int init(const lq_config* c) {
        device = c->device;
        LQ_CUDA(cudaSetDevice(device));
        if (c->n != N || c->m != M) LQ_FAIL(LQ_ERR_ARG, "model expects n=%d m=%d, config has n=%d m=%d", N, M, c->n, c->m);
        if (c->n_sys_params != Sys::NPARAMS) LQ_FAIL(LQ_ERR_ARG, "dynamics blob has %d entries, expected %d", c->n_sys_params, Sys::NPARAMS);
        if (c->horizon_iters < 1) LQ_FAIL(LQ_ERR_ARG, "horizon_iters must be >= 1");
        if (c->node_capacity < 1 || c->wave_capacity < 1) LQ_FAIL(LQ_ERR_ARG, "capacities must be >= 1");
        memset(&md, 0, sizeof md);
        md.sys.unpack(c->sys_params);
        if (md.lqr.unpack(c->lqr_params, c->n_lqr_params) != 0) LQ_FAIL(LQ_ERR_ARG, "lqr blob does not fit this provider/model");
        md.angle_mask = c->angle_mask;
        md.H = H = c->horizon_iters;
        md.dt = (T)c->dt;
        md.FPR = c->FPR;
        md.hfactor = c->hfactor; md.hlo = c->hspan_lo; md.hhi = c->hspan_hi;
        if (c->hfactor) {
            if (c->hfactor < 2 || c->hspan_lo < 1 || c->hspan_hi < c->hspan_lo || c->hspan_hi > c->horizon_iters)
                LQ_FAIL(LQ_ERR_ARG, "adaptive horizon needs hfactor >= 2 and 1 <= hspan_lo <= hspan_hi <= horizon_iters");
            if (c->hspan_hi > 127) LQ_FAIL(LQ_ERR_ARG, "adaptive horizon supports at most 127 steps");
        }
        for (int i = 0; i < N; i++) {
            md.tol[i] = (T)c->error_tol[i];
            md.goal_lo[i] = -Num<T>::inf(); md.goal_hi[i] = Num<T>::inf();
        }
        cap = (c->node_capacity + 1 + 3) & ~3;      // one spare row for the finish_on_goal node (planner.py:299)
        Wc = (c->wave_capacity + 3) & ~3;
        tries_cap = std::min(16, std::max(1, c->tries_capacity));      // the sampler evaluates tries on 16 lanes

        // feasibility tables
        if (c->n_prims < 0 || c->n_prims > LQ_MAX_PRIMS) LQ_FAIL(LQ_ERR_ARG, "at most %d feasibility primitives", LQ_MAX_PRIMS);
        md.feas.n_prims = c->n_prims;
        for (int pi = 0; pi < c->n_prims; pi++) {
            const lq_feas_prim& s = c->prims[pi];
            FeasPrim<T>& d = md.feas.prims[pi];
            d.kind = s.kind; d.n_obs = s.n_obs; d.n_verts = s.n_verts; d.flags = s.flags;
            for (int k = 0; k < 4; k++) d.dims[k] = s.dims[k];
            d.inflate = (T)s.inflate; d.bound_r = T(0); d.obs = nullptr; d.verts = nullptr; d.nbox = 0;
            const bool is_box = s.kind == LQ_FEAS_STATE_BOX || s.kind == LQ_FEAS_STATE_REGION || s.kind == LQ_FEAS_EFFORT_BOX;
            if (is_box) {
                d.nbox = s.kind == LQ_FEAS_EFFORT_BOX ? M : N;
                if (!s.lo || !s.hi) LQ_FAIL(LQ_ERR_ARG, "box primitive %d needs lo/hi", pi);
                for (int k = 0; k < LQ_MAX_N; k++) { d.lo[k] = -Num<T>::inf(); d.hi[k] = Num<T>::inf(); }   // dims beyond nbox never trip
                for (int k = 0; k < d.nbox; k++) { d.lo[k] = (T)s.lo[k]; d.hi[k] = (T)s.hi[k]; }
                continue;
            }
            if (s.kind == LQ_FEAS_OGRID_CLOUD) {
                if (!s.lo || !s.obs || !s.verts || s.n_obs < 1 || s.n_verts < 1 || s.lo[0] < 1) LQ_FAIL(LQ_ERR_ARG, "occupancy-grid primitive %d needs grid, vertices and {cols, cpm, ox, oy, threshold}", pi);
                for (int k = 0; k < 3; k++)
                    if (s.dims[k] < 0 || s.dims[k] >= N) LQ_FAIL(LQ_ERR_ARG, "primitive %d: state dim %d out of range", pi, s.dims[k]);
                const size_t cells = (size_t)s.n_obs * (size_t)s.lo[0];
                std::vector<T> grid(cells), vv((size_t)2 * s.n_verts);
                for (size_t k = 0; k < cells; k++) grid[k] = (T)s.obs[k];
                for (int k = 0; k < 2 * s.n_verts; k++) vv[k] = (T)s.verts[k];
                auto* gb = new DevBuf<T>(); feas_bufs.push_back(gb);
                auto* vb = new DevBuf<T>(); feas_bufs.push_back(vb);
                if (int r = upload(*gb, grid)) return r;
                if (int r = upload(*vb, vv)) return r;
                d.obs = gb->p; d.verts = vb->p; d.cell_start = nullptr; d.cell_items = nullptr; d.hull_axes = nullptr;
                for (int k = 0; k < 5; k++) d.lo[k] = (T)s.lo[k];
                continue;
            }
            int stride;
            switch (s.kind) {
            case LQ_FEAS_CIRCLES_POINT: case LQ_FEAS_CIRCLES_CLOUD: stride = 3; break;
            case LQ_FEAS_BOXES_CLOUD: case LQ_FEAS_SPHERES_POINT: stride = 4; break;
            default: LQ_FAIL(LQ_ERR_ARG, "unknown feasibility primitive kind %d", s.kind);
            }
            const int nd = s.kind == LQ_FEAS_SPHERES_POINT ? 3 : (s.kind == LQ_FEAS_CIRCLES_POINT ? 2 : 3);
            for (int k = 0; k < nd; k++)
                if (s.dims[k] < 0 || s.dims[k] >= N) LQ_FAIL(LQ_ERR_ARG, "primitive %d: state dim %d out of range", pi, s.dims[k]);
            std::vector<T> obs((size_t)std::max(1, s.n_obs) * 4, T(0));
            for (int o = 0; o < s.n_obs; o++) {
                for (int k = 0; k < stride; k++) obs[o * 4 + k] = (T)s.obs[o * stride + k];
                // point primitives compare against inflate + r (demo_boat_novice.py:162): fold it in float64
                if (s.kind == LQ_FEAS_CIRCLES_POINT) obs[o * 4 + 2] = (T)(s.inflate + s.obs[o * stride + 2]);
                if (s.kind == LQ_FEAS_SPHERES_POINT) obs[o * 4 + 3] = (T)(s.inflate + s.obs[o * stride + 3]);
            }
            auto* ob = new DevBuf<T>(); feas_bufs.push_back(ob);
            if (int r = upload(*ob, obs)) return r;
            d.obs = ob->p;
            d.cell_start = nullptr; d.cell_items = nullptr; d.n_cells = d.n_items = 0;
            d.hull_axes = nullptr; d.hull_nx = d.hull_ny = 0;
            if (s.kind == LQ_FEAS_CIRCLES_CLOUD || s.kind == LQ_FEAS_BOXES_CLOUD) {
                if (s.n_verts < 1 || !s.verts) LQ_FAIL(LQ_ERR_ARG, "cloud primitive %d needs vertices", pi);
                std::vector<T> vv((size_t)2 * s.n_verts);
                double br = 0;
                for (int v = 0; v < s.n_verts; v++) {
                    vv[v] = (T)s.verts[v]; vv[s.n_verts + v] = (T)s.verts[s.n_verts + v];
                    br = std::max(br, std::hypot(s.verts[v], s.verts[s.n_verts + v]));
                }
                d.bound_r = (T)br;
                auto* vb = new DevBuf<T>(); feas_bufs.push_back(vb);
                if (int r = upload(*vb, vv)) return r;
                d.verts = vb->p;
                // product-grid hull? (sorted unique x values) x (sorted unique y values) == the vertex set
                std::vector<double> ux(s.verts, s.verts + s.n_verts), uy(s.verts + s.n_verts, s.verts + 2 * s.n_verts);
                std::sort(ux.begin(), ux.end()); ux.erase(std::unique(ux.begin(), ux.end()), ux.end());
                std::sort(uy.begin(), uy.end()); uy.erase(std::unique(uy.begin(), uy.end()), uy.end());
                if (use_hull_grid && ux.size() * uy.size() == (size_t)s.n_verts && ux.size() + uy.size() < (size_t)s.n_verts) {
                    std::vector<char> seen((size_t)s.n_verts, 0);
                    bool full = true;
                    for (int v = 0; v < s.n_verts && full; v++) {
                        const size_t i = std::lower_bound(ux.begin(), ux.end(), s.verts[v]) - ux.begin();
                        const size_t j = std::lower_bound(uy.begin(), uy.end(), s.verts[s.n_verts + v]) - uy.begin();
                        if (seen[i * uy.size() + j]) full = false;
                        seen[i * uy.size() + j] = 1;
                    }
                    if (full) {
                        std::vector<T> ax;
                        for (double v : ux) ax.push_back((T)v);
                        for (double v : uy) ax.push_back((T)v);
                        auto* hb = new DevBuf<T>(); feas_bufs.push_back(hb);
                        if (int r = upload(*hb, ax)) return r;
                        d.hull_axes = hb->p; d.hull_nx = (int)ux.size(); d.hull_ny = (int)uy.size();
                    }
                }
            }
            if (s.n_obs >= grid_min_obs)
                if (int r = build_grid(s, d)) return r;
        }

        // shared-memory staging image for the rollout kernels: all tables of the small obstacle fields, as long as
        // they fit next to four resident CTAs per SM, laid out as the kernels read them (models.cuh stage_tables)
        if (int r = build_stage_image()) return r;

        // tree
        LQ_CUDA(t_state.alloc((size_t)N * cap));
        LQ_CUDA(t_K.alloc((size_t)M * N * cap));
        LQ_CUDA(t_ex.alloc((size_t)cap * H * N));
        LQ_CUDA(t_eu.alloc((size_t)cap * H * M));
        LQ_CUDA(t_parent.alloc(cap)); LQ_CUDA(t_cost.alloc(cap)); LQ_CUDA(t_elen.alloc(cap));
        LQ_CUDA(t_ignore.alloc((size_t)2 * cap)); LQ_CUDA(t_owner.alloc(cap)); LQ_CUDA(t_scal.alloc(8)); LQ_CUDA(t_best.alloc(1));
        LQ_CUDA(t_goals.alloc((size_t)2 * GOAL_CAP));
        tv.state = t_state.p; tv.K = t_K.p; tv.parent = t_parent.p; tv.cost = t_cost.p; tv.edge_len = t_elen.p;
        tv.ignore[0] = t_ignore.p; tv.ignore[1] = t_ignore.p + cap; tv.owner = t_owner.p; tv.edge_x = t_ex.p; tv.edge_u = t_eu.p;
        tv.size = t_scal.p; tv.n_ignored2[0] = t_scal.p + 1; tv.n_goal = t_scal.p + 2; tv.n_ignored2[1] = t_scal.p + 3;
        tv.goal_count[0] = t_scal.p + 4; tv.goal_count[1] = t_scal.p + 5;
        tv.goal_list[0] = t_goals.p; tv.goal_list[1] = t_goals.p + GOAL_CAP; tv.goal_cap = GOAL_CAP;
        tv.best = t_best.p; tv.cap = cap; tv.limit = c->node_capacity;
        LQ_CUDA(cudaMemset(t_scal.p, 0, 8 * sizeof(int)));
        LQ_CUDA(w_ctl.alloc(1));
        LQ_CUDA(cudaMemset(w_ctl.p, 0, sizeof(WaveCtl)));

        // wave working set; samples / tries / partial minima / per-query metrics / exchange segments are
        // double-buffered by wave parity: wave k+1's sampler and old-tree scan run while wave k still reads its own
        LQ_CUDA(w_samples.alloc((size_t)2 * N * Wc)); LQ_CUDA(w_ncost.alloc(Wc)); LQ_CUDA(w_nidx.alloc(Wc));
        LQ_CUDA(w_pcost.alloc((size_t)2 * MAX_SLOTS * Wc)); LQ_CUDA(w_pidx.alloc((size_t)2 * MAX_SLOTS * Wc));
        LQ_CUDA(w_xs.alloc((size_t)2 * Wc * H * N)); LQ_CUDA(w_us.alloc((size_t)2 * Wc * H * M));
        LQ_CUDA(w_len.alloc(Wc)); LQ_CUDA(w_tries.alloc((size_t)2 * Wc)); LQ_CUDA(w_pick.alloc(Wc));
        LQ_CUDA(w_misc.alloc(8)); LQ_CUDA(w_hor.alloc(1));
        LQ_CUDA(cudaMemset(w_misc.p, 0, 8 * sizeof(int)));      // [4] = completion counter of closest_kernel (zero between launches)
        { const int h0 = c->hfactor ? c->hspan_lo : c->horizon_iters; LQ_CUDA(cudaMemcpy(w_hor.p, &h0, sizeof(int), cudaMemcpyHostToDevice)); }
        md.horizon = w_hor.p;
        LQ_CUDA(w_goal.alloc(2 * LQ_MAX_N));
        seg_sz = seg_bytes<T>(Wc, N, M);
        LQ_CUDA(w_seg.alloc(2 * seg_sz));
        LQ_CUDA(cudaMemset(w_seg.p, 0, 2 * seg_sz));
        LQ_CUDA(cudaMemset(w_tries.p, 0, (size_t)2 * Wc * sizeof(int)));
        LQ_CUDA(cudaHostAlloc((void**)&h_ring, RING * sizeof(DevStats), cudaHostAllocMapped));
        LQ_CUDA(cudaHostGetDevicePointer((void**)&d_ring, h_ring, 0));
        LQ_CUDA(cudaHostAlloc((void**)&h_stop, 64, cudaHostAllocMapped));
        LQ_CUDA(cudaHostGetDevicePointer((void**)&d_stop, h_stop, 0));
        LQ_CUDA(cudaMallocHost((void**)&h_misc, 8 * sizeof(int)));
        memset(h_ring, 0, RING * sizeof(DevStats));
        *h_stop = 0;
        memset(&last_stats, 0, sizeof last_stats);

        // metric: constant providers are classified once (planner.py:348 evaluates lqr at every sample)
        if (Lqr::CONST_METRIC) {
            metric_diag = true;
            const T* S = const_metric_ptr();
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    Sconst[i * N + j] = (double)S[i * N + j];
                    if (i != j && S[i * N + j] != T(0)) metric_diag = false;
                }
            for (int i = 0; i < N; i++) if (Sconst[i * N + i] < 0) metric_diag = false;
        } else {
            metric_diag = false;
        }
        if (metric_diag) {
            dmet.angle_mask = md.angle_mask;
            dmet.exact = sizeof(T) == 8 ? 1 : 0;
            for (int i = 0; i < N; i++) {
                const double w = std::sqrt(Sconst[i * N + i]);
                dmet.w[i] = (T)w; dmet.period[i] = (T)(6.283185307179586 * w); dmet.sdiag[i] = (T)Sconst[i * N + i];
            }
        } else {
            LQ_CUDA(w_Sp.alloc((size_t)2 * NS * Wc));
            if (Lqr::CONST_METRIC) {
                std::vector<T> sp_h((size_t)NS * Wc);
                int s = N;
                for (int a = 0; a < N; a++) {
                    for (int j = 0; j < Wc; j++) sp_h[(size_t)a * Wc + j] = (T)Sconst[a * N + a];
                    for (int b = a + 1; b < N; b++, s++)
                        for (int j = 0; j < Wc; j++) sp_h[(size_t)s * Wc + j] = (T)(Sconst[a * N + b] + Sconst[b * N + a]);
                }
                for (int par = 0; par < 2; par++)
                    LQ_CUDA(cudaMemcpy(w_Sp.p + (size_t)par * NS * Wc, sp_h.data(), sp_h.size() * sizeof(T), cudaMemcpyHostToDevice));
            }
        }
        // heuristic: split the obstacle tables over a warp once they dominate a rollout step
        int obs_work = 0;
        for (int pi = 0; pi < md.feas.n_prims; pi++) obs_work += md.feas.prims[pi].n_obs;
        // sub-warp groups of 8 measured best on C4 (profiles/README.md): they split hull vertices / obstacle
        // lists and cut the divergence of thread-per-sample rollouts; without obstacle tables one lane is enough
        steer_lanes = obs_work > 0 ? 8 : 1;
        if (Lqr::COOP) steer_lanes = 32;       // a warp per sample shares the Riccati sweep (measured 1.7x over 8 lanes on C5)
        tv.zstate = nullptr; tv.U = nullptr;
        memset(&tv.grid, 0, sizeof tv.grid);
        grid_ok = metric_diag && N >= 2 && !(md.angle_mask & 3u) && Sconst[0] > 0 && Sconst[N + 1] > 0;
        if (grid_ok) {
            constexpr int FW = GRID_F * GRID_F / 32, CW = GRID_C * GRID_C / 32;
            LQ_CUDA(g_head.alloc((size_t)GRID_F * GRID_F)); LQ_CUDA(g_next.alloc(cap)); LQ_CUDA(g_bits.alloc(FW + CW));
            tv.grid.head = g_head.p; tv.grid.next = g_next.p; tv.grid.fine = g_bits.p; tv.grid.coarse = g_bits.p + FW;
            tv.grid.dx = 0; tv.grid.dy = 1; tv.grid.on = 0;         // linked and searched only while the option is on
            grid_extents(0.0, 1.0, 0.0, 1.0);
            grid_have_extents = false; grid_dirty = false;
        }
        return LQ_OK;
    }

    // ---- spatial index ----------------------------------------------------------
    // Any extents are correct (everything beyond the grid is filed under its border cells); they only decide how
    // well the cells resolve the tree.  Taken from the sample box (planner.py:194-198), else from the loaded nodes.
    void grid_extents(double lo0, double hi0, double lo1, double hi1) {
        if (!grid_ok) return;
        auto axis = [](double lo, double hi, float* origin, float* inv_h, double* h) {
            if (!(hi > lo) || !std::isfinite(lo) || !std::isfinite(hi)) { lo = lo - 0.5; hi = lo + 1.0; if (!std::isfinite(lo)) { lo = 0; hi = 1; } }
            const double pad = 0.02 * (hi - lo);
            lo -= pad; hi += pad;
            *h = (hi - lo) / GRID_F;
            *origin = (float)lo; *inv_h = (float)(1.0 / *h);
        };
        GridView g = tv.grid;
        double hx = 1, hy = 1;
        axis(lo0, hi0, &g.x0, &g.inv_hx, &hx);
        axis(lo1, hi1, &g.y0, &g.inv_hy, &hy);
        g.kx = (float)(hx * hx * Sconst[0]); g.ky = (float)(hy * hy * Sconst[N + 1]);
        if (g.x0 != tv.grid.x0 || g.y0 != tv.grid.y0 || g.inv_hx != tv.grid.inv_hx || g.inv_hy != tv.grid.inv_hy) grid_dirty = true;
        tv.grid = g;
        grid_have_extents = true;
    }
    // (re)link every node of the tree, on `st`
    int grid_rebuild(cudaStream_t st, int count_ub) {
        if (!grid_active()) return LQ_OK;
        LQ_CUDA(cudaMemsetAsync(g_head.p, 0xff, g_head.n * sizeof(int), st));
        LQ_CUDA(cudaMemsetAsync(g_next.p, 0xfe, g_next.n * sizeof(int), st));        // GRID_PENDING
        LQ_CUDA(cudaMemsetAsync(g_bits.p, 0, g_bits.n * sizeof(unsigned), st));
        const int n = std::max(1, count_ub);
        grid_insert_range_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(tv, 0, tv.size, n); launches++;
        LQ_CUDA(cudaGetLastError());
        grid_dirty = false;
        return LQ_OK;
    }
    int grid_sync(cudaStream_t st) { return (grid_active() && grid_dirty) ? grid_rebuild(st, size_ub) : LQ_OK; }
    bool grid_active() const { return grid_ok && grid_opt == 1; }

    // ---- state-independent metric of a state-dependent provider ---------------------
    bool whiten_active() const { return whiten_on && whiten_opt != 0; }
    T* zq_of(int par) { return w_zq.p + (size_t)par * N * Wc; }
    // Probe S = lqr(x, 0).S (planner.py:348) at the centre of the sample box and at points spread over it.  If every probe
    // agrees with their mean to within the finite-difference noise of the arithmetic (float32 differences at |x| ~ 50
    // are noisy at the 1e-3 level; float64 at 1e-11), the metric is that mean for every sample: factor it, keep z = U v.
    int whiten_probe(const double* lo, const double* hi) {
        if (Lqr::CONST_METRIC || md.angle_mask != 0u || whiten_opt == 0) return LQ_OK;
        bool same = whiten_probed;
        for (int i = 0; i < N && same; i++) same = whiten_box[2 * i] == lo[i] && whiten_box[2 * i + 1] == hi[i];
        if (same) return LQ_OK;
        for (int i = 0; i < N; i++) { whiten_box[2 * i] = lo[i]; whiten_box[2 * i + 1] = hi[i]; }
        whiten_probed = true;
        constexpr int P = 12;
        std::vector<double> x((size_t)P * N), u((size_t)P * M, 0.0), S((size_t)P * N * N), K((size_t)P * M * N);
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        for (int p = 0; p < P; p++)
            for (int i = 0; i < N; i++) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const double r = p == 0 ? 0.5 : (double)(rng >> 11) / 9007199254740992.0;
                const double a = std::isfinite(lo[i]) ? lo[i] : 0.0, b = std::isfinite(hi[i]) ? hi[i] : a;
                x[(size_t)p * N + i] = a + (b - a) * r;
            }
        if (int r = eval_lqr(P, x.data(), u.data(), S.data(), K.data())) return r;
        double mean[N * N];
        double scale = 0, dev = 0;
        for (int k = 0; k < N * N; k++) {
            double m = 0;
            for (int p = 0; p < P; p++) m += S[(size_t)p * N * N + k];
            mean[k] = m / P;
        }
        for (int a = 0; a < N; a++)
            for (int b = a + 1; b < N; b++) mean[a * N + b] = mean[b * N + a] = 0.5 * (mean[a * N + b] + mean[b * N + a]);
        for (int k = 0; k < N * N; k++) {
            if (!std::isfinite(mean[k])) { whiten_on = false; tv.zstate = nullptr; return LQ_OK; }
            scale = std::max(scale, std::fabs(mean[k]));
            for (int p = 0; p < P; p++) dev = std::max(dev, std::fabs(S[(size_t)p * N * N + k] - mean[k]));
        }
        const double tol = sizeof(T) == 4 ? 3e-2 : 1e-7;
        bool ok = scale > 0 && dev <= tol * scale;
        // Cholesky S = C C^T; U = C^T (upper triangular), so that e^T S e = |U e|^2
        double C[N * N] = {0};
        for (int i = 0; i < N && ok; i++)
            for (int j = 0; j <= i; j++) {
                double a = mean[i * N + j];
                for (int k = 0; k < j; k++) a -= C[i * N + k] * C[j * N + k];
                if (i == j) { if (!(a > 1e-12 * scale)) { ok = false; break; } C[i * N + i] = std::sqrt(a); }
                else C[i * N + j] = a / C[j * N + j];
            }
        if (!ok) { whiten_on = false; tv.zstate = nullptr; tv.U = nullptr; return LQ_OK; }
        std::vector<T> Uh((size_t)N * N, T(0));
        for (int i = 0; i < N; i++)
            for (int j = i; j < N; j++) Uh[(size_t)i * N + j] = (T)C[j * N + i];
        if (!t_z.p) { LQ_CUDA(t_z.alloc((size_t)N * cap)); LQ_CUDA(w_zq.alloc((size_t)2 * N * Wc)); }
        if (int r = upload(w_U, Uh)) return r;
        for (int i = 0; i < N; i++) { dmet_w.w[i] = T(1); dmet_w.period[i] = (T)6.283185307179586; dmet_w.sdiag[i] = T(1); }
        dmet_w.angle_mask = 0u; dmet_w.exact = sizeof(T) == 8 ? 1 : 0;
        for (int k = 0; k < N * N; k++) Sconst[k] = mean[k];
        tv.zstate = t_z.p; tv.U = w_U.p;
        whiten_on = true; z_dirty = true;
        return LQ_OK;
    }
    int whiten_range(int first, const int* d_size, int count_ub, cudaStream_t st) {
        const int n = std::max(1, count_ub);
        whiten_range_kernel<T, N><<<(n + 127) / 128, 128, 0, st>>>(t_state.p, t_z.p, w_U.p, cap, first, d_size, n); launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }
    int whiten_sync(cudaStream_t st) {
        if (!whiten_on || !z_dirty) return LQ_OK;
        z_dirty = false;
        return whiten_range(0, tv.size, size_ub, st);
    }

    // Uniform grid over the obstacle table (see FeasPrim).  Zones of influence are axis-aligned boxes:
    // centre +- (obstacle extent + hull bounding radius + margin); the margin absorbs float rounding of
    // the cell index on the device.
    int build_grid(const lq_feas_prim& s, FeasPrim<T>& d) {
        const int nd = s.kind == LQ_FEAS_SPHERES_POINT ? 3 : 2;
        const int stride = (s.kind == LQ_FEAS_CIRCLES_POINT || s.kind == LQ_FEAS_CIRCLES_CLOUD) ? 3 : 4;
        const int O = s.n_obs;
        std::vector<double> lo((size_t)O * 3, 0.0), hi((size_t)O * 3, 0.0);
        double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300}, mean_ext = 0;
        for (int o = 0; o < O; o++) {
            const double* ob = s.obs + (size_t)o * stride;
            double ext[3];
            switch (s.kind) {
            case LQ_FEAS_CIRCLES_POINT: ext[0] = ext[1] = s.inflate + ob[2]; ext[2] = 0; break;
            case LQ_FEAS_SPHERES_POINT: ext[0] = ext[1] = ext[2] = s.inflate + ob[3]; break;
            case LQ_FEAS_CIRCLES_CLOUD: ext[0] = ext[1] = ob[2] + (double)d.bound_r; ext[2] = 0; break;
            default: ext[0] = ob[2] + (double)d.bound_r; ext[1] = ob[3] + (double)d.bound_r; ext[2] = 0; break;
            }
            for (int k = 0; k < nd; k++) {
                const double e = std::max(0.0, ext[k]) * 1.001 + 1e-3 * (1.0 + std::fabs(ob[k]));
                lo[o * 3 + k] = ob[k] - e; hi[o * 3 + k] = ob[k] + e;
                blo[k] = std::min(blo[k], lo[o * 3 + k]); bhi[k] = std::max(bhi[k], hi[o * 3 + k]);
                mean_ext += e / (nd * O);
            }
        }
        int gn[3] = {1, 1, 1};
        double cell[3] = {1, 1, 1};
        for (int k = 0; k < nd; k++) {
            const double span = std::max(1e-9, bhi[k] - blo[k]);
            gn[k] = (int)std::min(nd == 3 ? 48.0 : 256.0, std::max(1.0, std::ceil(span / (2.0 * mean_ext))));
            cell[k] = span / gn[k] * (1.0 + 1e-9);
        }
        const size_t ncell = (size_t)gn[0] * gn[1] * gn[2];
        std::vector<int> count(ncell + 1, 0);
        auto range = [&](int o, int k, int& a, int& b) {
            a = std::max(0, std::min(gn[k] - 1, (int)std::floor((lo[o * 3 + k] - blo[k]) / cell[k])));
            b = std::max(0, std::min(gn[k] - 1, (int)std::floor((hi[o * 3 + k] - blo[k]) / cell[k])));
        };
        for (int pass = 0; pass < 2; pass++) {
            std::vector<int> fill;
            if (pass == 1) {
                for (size_t c = 0, run = 0; c <= ncell; c++) { const int n = count[c]; count[c] = (int)run; run += n; }
                fill.assign(count.begin(), count.end());
                items_tmp.assign(count[ncell], 0);
            }
            for (int o = 0; o < O; o++) {
                int a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
                for (int k = 0; k < nd; k++) range(o, k, a[k], b[k]);
                for (int z = a[2]; z <= b[2]; z++)
                    for (int y = a[1]; y <= b[1]; y++)
                        for (int x = a[0]; x <= b[0]; x++) {
                            const size_t c = ((size_t)z * gn[1] + y) * gn[0] + x;
                            if (pass == 0) count[c]++; else items_tmp[fill[c]++] = o;
                        }
            }
        }
        auto* cs = new DevBuf<int>(); grid_bufs.push_back(cs);
        auto* ci = new DevBuf<int>(); grid_bufs.push_back(ci);
        if (int r = upload(*cs, count)) return r;
        if (items_tmp.empty()) items_tmp.push_back(0);
        if (int r = upload(*ci, items_tmp)) return r;
        d.cell_start = cs->p; d.cell_items = ci->p;
        d.n_cells = (int)ncell; d.n_items = (int)items_tmp.size();
        for (int k = 0; k < 3; k++) { d.g0[k] = (T)blo[k]; d.ginv[k] = (T)(1.0 / cell[k]); d.gn[k] = gn[k]; }
        return LQ_OK;
    }
    std::vector<int> items_tmp;
    DevBuf<unsigned char> stage_img;

    int build_stage_image() {
        memset(&stage_plan, 0, sizeof stage_plan);
        for (int pi = 0; pi < LQ_MAX_PRIMS; pi++) stage_plan.off_obs[pi] = stage_plan.off_cs[pi] = stage_plan.off_ci[pi] = stage_plan.off_hull[pi] = -1;
        std::vector<unsigned char> img;
        auto put = [&](const void* src, size_t bytes) -> int {
            const int off = (int)img.size();
            img.resize(off + stage_align((int)bytes), 0);
            memcpy(img.data() + off, src, bytes);
            return off;
        };
        for (int pi = 0; pi < md.feas.n_prims; pi++) {
            const FeasPrim<T>& d = md.feas.prims[pi];
            const bool tabular = d.kind == LQ_FEAS_CIRCLES_POINT || d.kind == LQ_FEAS_SPHERES_POINT ||
                                 d.kind == LQ_FEAS_CIRCLES_CLOUD || d.kind == LQ_FEAS_BOXES_CLOUD;
            if (!tabular || d.n_obs < 1) continue;
            int need = stage_align(d.n_obs * 4 * (int)sizeof(T));
            if (d.cell_start) need += stage_align((d.n_cells + 1) * (int)sizeof(int)) + stage_align(d.n_items * (int)sizeof(int));
            if (d.hull_axes) need += stage_align((d.hull_nx + d.hull_ny) * (int)sizeof(T));
            if ((int)img.size() + need > 40 * 1024) continue;
            // read the tables back from the device: they are what the un-staged path would read
            std::vector<T> obs((size_t)d.n_obs * 4);
            LQ_CUDA(cudaMemcpy(obs.data(), d.obs, obs.size() * sizeof(T), cudaMemcpyDeviceToHost));
            stage_plan.off_obs[pi] = put(obs.data(), obs.size() * sizeof(T));
            if (d.cell_start) {
                std::vector<int> cs((size_t)d.n_cells + 1), ci((size_t)std::max(1, d.n_items));
                LQ_CUDA(cudaMemcpy(cs.data(), d.cell_start, cs.size() * sizeof(int), cudaMemcpyDeviceToHost));
                LQ_CUDA(cudaMemcpy(ci.data(), d.cell_items, (size_t)d.n_items * sizeof(int), cudaMemcpyDeviceToHost));
                stage_plan.off_cs[pi] = put(cs.data(), cs.size() * sizeof(int));
                stage_plan.off_ci[pi] = put(ci.data(), (size_t)d.n_items * sizeof(int));
            }
            if (d.hull_axes) {
                std::vector<T> hx((size_t)d.hull_nx + d.hull_ny);
                LQ_CUDA(cudaMemcpy(hx.data(), d.hull_axes, hx.size() * sizeof(T), cudaMemcpyDeviceToHost));
                stage_plan.off_hull[pi] = put(hx.data(), hx.size() * sizeof(T));
            }
            stage_plan.on[pi] = 1;
        }
        stage_plan.bytes = (int)img.size();
        if (!img.empty()) {
            if (int r = upload(stage_img, img)) return r;
            stage_plan.image = stage_img.p;
        }
        return LQ_OK;
    }

    const T* const_metric_ptr() { return const_metric_ptr_impl(md.lqr, 0); }
    template <class L> static auto const_metric_ptr_impl(L& l, int) -> decltype((const T*)l.S) { return l.S; }
    template <class L> static const T* const_metric_ptr_impl(L&, long) { return nullptr; }

    // ---- problem state --------------------------------------------------------
    int set_goal(const double* goal, const double* buf) override {
        LQ_CUDA(cudaSetDevice(device));
        std::vector<T> g(2 * LQ_MAX_N, T(0));
        for (int i = 0; i < N; i++) {
            // goal region (goal - buff, goal + buff), planner.py:482-484; evaluated in float64 first
            md.goal_lo[i] = (T)(goal[i] - buf[i]); md.goal_hi[i] = (T)(goal[i] + buf[i]);
            g[i] = md.goal_lo[i]; g[LQ_MAX_N + i] = md.goal_hi[i];
            sp.goal[i] = (T)goal[i];
        }
        LQ_CUDA(cudaMemcpy(w_goal.p, g.data(), g.size() * sizeof(T), cudaMemcpyHostToDevice));
        have_goal = true;
        return LQ_OK;
    }

    int set_horizon(int horizon_iters, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (!md.hfactor) LQ_FAIL(LQ_ERR_STATE, "the engine was built with a fixed horizon");
        if (horizon_iters < md.hlo || horizon_iters > md.hhi) LQ_FAIL(LQ_ERR_ARG, "horizon_iters %d outside [%d, %d]", horizon_iters, md.hlo, md.hhi);
        LQ_CUDA(cudaMemcpyAsync(w_hor.p, &horizon_iters, sizeof(int), cudaMemcpyHostToDevice, st));
        LQ_CUDA(cudaStreamSynchronize(st));
        return LQ_OK;
    }

    int set_sampling(const double* lo, const double* hi, const double* bias, int tries) override {
        if (tries < 1 || tries > tries_cap) LQ_FAIL(LQ_ERR_ARG, "tries must be in 1..%d", tries_cap);
        for (int i = 0; i < N; i++) {
            sp.center[i] = (T)((lo[i] + hi[i]) / 2.0);      // np.mean(sample_space, axis=1), planner.py:197
            sp.span[i] = (T)(hi[i] - lo[i]);                 // np.diff, planner.py:198
            sp.bias[i] = (T)bias[i];
        }
        sp.tries = tries;
        have_sampling = true;
        if (N >= 2) grid_extents(lo[0], hi[0], lo[1], hi[1]);
        return whiten_probe(lo, hi);
    }

    // New tree = the seed alone (planner.py:172-173, tree.py:50-73).  The seed state rides in the kernel arguments:
    // nothing is staged on the host, nothing waits -- everything is ordered on `st`, which the wave pipeline picks up.
    int reset(const double* x0, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        SeedArgs<T, N> sa;
        for (int d = 0; d < N; d++) sa.x[d] = (T)x0[d];
        LQ_CUDA(cudaMemsetAsync(t_ignore.p, 0, (size_t)2 * cap, st)); ign_copy = 0;
        LQ_CUDA(cudaMemsetAsync(t_owner.p, 0, cap, st));
        seed_kernel<T, N><<<1, 32, 0, st>>>(tv, sa); launches++;
        tree_scalars_kernel<T><<<1, 1, 0, st>>>(tv, 1, w_ctl.p); launches++;
        init_gain_kernel<T, Sys, Lqr><<<1, 128, 0, st>>>(md, t_state.p, t_K.p, cap, 1); launches++;
        LQ_CUDA(cudaGetLastError());
        if (whiten_on) { if (int r = whiten_range(0, nullptr, 1, st)) return r; z_dirty = false; }
        if (grid_ok) {
            if (!grid_have_extents) { grid_extents(x0[0] - 50.0, x0[0] + 50.0, x0[1] - 50.0, x0[1] + 50.0); grid_have_extents = false; }
            if (int r = grid_rebuild(st, 1)) return r;
            if (!grid_active()) grid_dirty = true;
        }
        size_ub = 1; size_ub_exact = 1;
        memset(&last_stats, 0, sizeof last_stats);
        last_stats.tree_size = 1; last_stats.best_node = -1; last_stats.best_cost = -1;
        q_head = q_tail = 0;
        for (int k = 0; k < RING; k++) prof_open[k] = false;
        sample_parity = 0; consumed_valid[0] = consumed_valid[1] = false;
        return LQ_OK;
    }

    int load_nodes(int count, const double* state, const int* parent, const int* cost, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (count < 1 || count > tv.limit) LQ_FAIL(LQ_ERR_ARG, "node count %d outside 1..%d", count, tv.limit);
        std::vector<T> soa((size_t)N * count);
        for (int i = 0; i < count; i++)
            for (int d = 0; d < N; d++) soa[(size_t)d * count + i] = (T)state[(size_t)i * N + d];
        LQ_CUDA(cudaStreamSynchronize(st));
        for (int d = 0; d < N; d++)
            LQ_CUDA(cudaMemcpyAsync(t_state.p + (size_t)d * cap, soa.data() + (size_t)d * count, count * sizeof(T),
                                    cudaMemcpyHostToDevice, st));
        std::vector<int> par(count, -1), cst(count, 0);
        if (parent) par.assign(parent, parent + count);
        if (cost) cst.assign(cost, cost + count);
        LQ_CUDA(cudaMemcpyAsync(t_parent.p, par.data(), count * sizeof(int), cudaMemcpyHostToDevice, st));
        LQ_CUDA(cudaMemcpyAsync(t_cost.p, cst.data(), count * sizeof(int), cudaMemcpyHostToDevice, st));
        LQ_CUDA(cudaMemsetAsync(t_elen.p, 0, count * sizeof(int), st));
        LQ_CUDA(cudaMemsetAsync(t_ignore.p, 0, (size_t)2 * cap, st)); ign_copy = 0;
        LQ_CUDA(cudaMemsetAsync(t_owner.p, 0, cap, st));
        tree_scalars_kernel<T><<<1, 1, 0, st>>>(tv, count, w_ctl.p); launches++;
        init_gain_kernel<T, Sys, Lqr><<<(count + 127) / 128, 128, 0, st>>>(md, t_state.p, t_K.p, cap, count); launches++;
        LQ_CUDA(cudaGetLastError());
        if (whiten_on) { if (int r = whiten_range(0, nullptr, count, st)) return r; z_dirty = false; }
        if (grid_ok) {
            if (!grid_have_extents) {                 // no sample box yet: the loaded nodes' own bounding box
                double b[4] = {state[0], state[0], state[1], state[1]};
                for (int i = 0; i < count; i++) {
                    const double x = state[(size_t)i * N], y = state[(size_t)i * N + 1];
                    if (std::isfinite(x)) { b[0] = std::min(b[0], x); b[1] = std::max(b[1], x); }
                    if (std::isfinite(y)) { b[2] = std::min(b[2], y); b[3] = std::max(b[3], y); }
                }
                grid_extents(b[0], b[1], b[2], b[3]);
                grid_have_extents = false;            // a later sample box still takes over
            }
            if (int r = grid_rebuild(st, count)) return r;
            if (!grid_active()) grid_dirty = true;       // not linked: switching the option on re-links first
        }
        LQ_CUDA(cudaStreamSynchronize(st));     // host staging vectors die here
        size_ub = count; size_ub_exact = count;
        memset(&last_stats, 0, sizeof last_stats);
        last_stats.tree_size = count; last_stats.best_node = -1; last_stats.best_cost = -1;
        q_head = q_tail = 0;
        for (int k = 0; k < RING; k++) prof_open[k] = false;
        sample_parity = 0; consumed_valid[0] = consumed_valid[1] = false;
        return LQ_OK;
    }

    int tree_size(int* out) override {
        LQ_CUDA(cudaSetDevice(device));
        LQ_CUDA(cudaDeviceSynchronize());
        LQ_CUDA(cudaMemcpy(out, tv.size, sizeof(int), cudaMemcpyDeviceToHost));
        size_ub = *out;
        return LQ_OK;
    }

    // ---- K1 -------------------------------------------------------------------
    // One scan launch over the node range [*d_lo, *d_hi) (d_lo == nullptr: from 0) for the first *d_count (or W)
    // queries of `q`; per-split partial minima go to slots [slot0, slot0 + nsplit) of parity `par`'s partials.
    // W, n_ub are host-side bounds used only to shape the grid; returns the number of slots written.
    T* pcost_of(int par) { return w_pcost.p + (size_t)par * MAX_SLOTS * Wc; }
    int* pidx_of(int par) { return w_pidx.p + (size_t)par * MAX_SLOTS * Wc; }
    T* sp_of(int par) { return w_Sp.p ? w_Sp.p + (size_t)par * NS * Wc : nullptr; }
    T* samples_of(int par) { return w_samples.p + (size_t)par * N * Wc; }
    int* tries_of(int par) { return w_tries.p + (size_t)par * Wc; }
    char* seg_of(int par) { return w_seg.p + (size_t)par * seg_sz; }
    T* xs_of(int par) { return w_xs.p + (size_t)par * Wc * H * N; }
    T* us_of(int par) { return w_us.p + (size_t)par * Wc * H * M; }

    // Slots: [slot_base, slot_base + nsplit), or -- end_aligned -- [slot_base - nsplit, slot_base), so that an old-tree
    // scan and a delta scan of the same wave leave one contiguous range around MAX_SPLIT_OLD for the rollout kernel.
    int launch_scan(const T* q, int qstride, int W, const int* d_count, const int* d_lo, const int* d_hi, int n_ub,
                    const int* d_nign, int use_ignore, int par, int slot_base, bool end_aligned, int max_split, int target_ctas,
                    cudaStream_t st, int* nsplit_out, int copy = 0) {
        const uint8_t* ignp = tv.ignore[copy];
        const int Nub = std::max(1, n_ub);
        const int max_splits = std::max(1, std::min(max_split, (Nub + NEAR_BLOCK - 1) / NEAR_BLOCK));
        ScanArgs sa;
        sa.d_lo = d_lo; sa.d_hi = d_hi; sa.d_count = d_count; sa.d_nignored = d_nign; sa.W = W; sa.slot0 = slot_base;
        sa.use_ignore = use_ignore;
        auto plan = [&](int qper) {
            sa.nqb = std::max(1, (W + NEAR_BLOCK * qper - 1) / (NEAR_BLOCK * qper));
            sa.nsplit = std::max(1, std::min(max_splits, (target_ctas + sa.nqb - 1) / sa.nqb));
            sa.slot0 = end_aligned ? slot_base - sa.nsplit : slot_base;
        };
        T* pc = pcost_of(par); int* pi = pidx_of(par);
        // the erf masks of the shipped vehicles get compile-time kernels; anything else the generic one
        const int cmask = md.angle_mask == 0u ? 0 : (md.angle_mask == 4u && N >= 3 ? 4 : (md.angle_mask == 1u ? 1 : -1));
        if (metric_diag && d_lo == nullptr && grid_active() && !grid_dirty &&
            (long long)std::max(1, W) * Nub >= (long long)grid_min_kpairs * 1024) {
            // indexed search (grid.cuh): one candidate per query in a single slot
            constexpr bool EX = sizeof(T) == 8;
            constexpr int GL = GRID_R, GB = 128;
            sa.nqb = std::max(1, (W * GL + GB - 1) / GB); sa.nsplit = 1;
            sa.slot0 = end_aligned ? slot_base - 1 : slot_base;
#define LQ_NEAR_GRID(MM)                                                                                   \
    nearest_grid_kernel<T, N, MM, EX, GL, GB><<<sa.nqb, GB, 0, st>>>(t_state.p, cap, ignp, sa, q, qstride, dmet, tv.grid, pc, pi, Wc)
            if (cmask == 0) LQ_NEAR_GRID(0);
            else if (cmask == 4) LQ_NEAR_GRID((N >= 3 ? 4 : 0));
            else if (cmask == 1) LQ_NEAR_GRID(1);
            else LQ_NEAR_GRID(-1);
#undef LQ_NEAR_GRID
            launches++;
            if (nsplit_out) *nsplit_out = 1;
            LQ_CUDA(cudaGetLastError());
            return LQ_OK;
        }
        // a state-independent dense metric goes through the diagonal scans on the whitened copy: queries first
        const bool whitened = !metric_diag && whiten_active() && !z_dirty;
        const T* scan_state = t_state.p;
        const DiagMetric<T, N>& dmet = whitened ? dmet_w : this->dmet;
        if (whitened) {
            whiten_queries_kernel<T, N><<<(std::max(1, W) + 127) / 128, 128, 0, st>>>(q, qstride, W, d_count, w_U.p, zq_of(par), Wc);
            launches++;
            LQ_CUDA(cudaGetLastError());
            q = zq_of(par); qstride = Wc; scan_state = t_z.p;
        }
        if (metric_diag || whitened) {
            constexpr int QBIG = sizeof(T) == 4 ? 4 : 2;
            constexpr bool EX = sizeof(T) == 8;
            plan(QBIG);
            bool big = (long long)sa.nqb * max_splits >= 148 * 2;
            if (near_q == 1) big = false;
            if (near_q > 1) big = true;
            if (!big) plan(1);
#define LQ_NEAR_DIAG(QQ, MM)                                                                              \
    nearest_diag_kernel<T, N, QQ, NEAR_BLOCK, MM, EX><<<sa.nqb * sa.nsplit, NEAR_BLOCK, 0, st>>>(         \
        scan_state, cap, ignp, sa, q, qstride, dmet, pc, pi, Wc)
#define LQ_NEAR_DIAG_M(QQ)                                  \
    do {                                                    \
        if (cmask == 0) LQ_NEAR_DIAG(QQ, 0);                \
        else if (cmask == 4) LQ_NEAR_DIAG(QQ, (N >= 3 ? 4 : 0)); \
        else if (cmask == 1) LQ_NEAR_DIAG(QQ, 1);           \
        else LQ_NEAR_DIAG(QQ, -1);                          \
    } while (0)
#define LQ_NEAR_PK(QQ, MB, PK, MM)                                                                         \
    nearest_diag_pk_kernel<N, QQ, NEAR_BLOCK, MM, 8, MB, PK><<<sa.nqb * sa.nsplit, NEAR_BLOCK, 0, st>>>(    \
        scan_state, cap, ignp, sa, q, qstride, dmet, pc, pi, Wc)
#define LQ_NEAR_PK_M(QQ, MB, PK)                                  \
    do {                                                          \
        plan(QQ);                                                 \
        if (cmask == 0) LQ_NEAR_PK(QQ, MB, PK, 0);                \
        else if (cmask == 4) LQ_NEAR_PK(QQ, MB, PK, (N >= 3 ? 4 : 0)); \
        else if (cmask == 1) LQ_NEAR_PK(QQ, MB, PK, 1);           \
        else LQ_NEAR_PK(QQ, MB, PK, -1);                          \
    } while (0)
            bool packed = false;
            if constexpr (sizeof(T) == 4) {          // FP32x2 issue + two-level argmin: same results, fewer issue slots
                if (big && near_kernel != LQ_NEAREST_SCALAR) {
                    packed = true;
                    switch (near_kernel) {
                    // minimum CTAs per SM = the register budget each variant is compiled for (128-thread CTAs)
                    case LQ_NEAREST_PACKED: LQ_NEAR_PK_M(4, (N <= 6 ? 8 : 4), true); break;
                    case LQ_NEAREST_PACKED_Q8: LQ_NEAR_PK_M(8, (N <= 6 ? 4 : 2), true); break;
                    case LQ_NEAREST_PREFIX:                // partial-distance filter on the leading (position) dims
                        if constexpr (N >= 3) {
                            constexpr int PFX = N >= 12 ? 3 : 2;          // leading position dims used as the bound
                            if ((md.angle_mask & ((1u << PFX) - 1u)) == 0u && (cmask == 0 || cmask == 4)) {
                                plan(4);
#define LQ_NEAR_PFX(MM)                                                                                   \
    nearest_diag_prefix_kernel<N, 4, NEAR_BLOCK, MM, PFX><<<sa.nqb * sa.nsplit, NEAR_BLOCK, 0, st>>>(       \
        scan_state, cap, ignp, sa, q, qstride, dmet, pc, pi, Wc)
                                if (cmask == 0) LQ_NEAR_PFX(0); else LQ_NEAR_PFX((N >= 3 ? 4 : 0));
#undef LQ_NEAR_PFX
                                break;
                            }
                        }
                        LQ_NEAR_PK_M(4, (N <= 6 ? 8 : 4), true); break;      // not applicable: the packed scan
                    default: LQ_NEAR_PK_M(4, (N <= 6 ? 8 : 4), false); break;     // LQ_NEAREST_TWO_LEVEL
                    }
                }
            }
            if (!packed) { if (big) LQ_NEAR_DIAG_M(QBIG); else LQ_NEAR_DIAG_M(1); }
#undef LQ_NEAR_PK_M
#undef LQ_NEAR_PK
#undef LQ_NEAR_DIAG_M
#undef LQ_NEAR_DIAG
        } else {
            if (qstride != Wc) LQ_FAIL(LQ_ERR_STATE, "dense-metric search needs queries in the engine's wave layout");
            plan(1);
            const T* Sp = sp_of(par);
#define LQ_NEAR_DENSE(MM)                                                                                  \
    nearest_dense_kernel<T, N, NEAR_BLOCK, MM><<<sa.nqb * sa.nsplit, NEAR_BLOCK, 0, st>>>(                 \
        t_state.p, cap, ignp, sa, q, qstride, Sp, md.angle_mask, pc, pi, Wc)
            if (cmask == 0) LQ_NEAR_DENSE(0);
            else if (cmask == 4) LQ_NEAR_DENSE((N >= 3 ? 4 : 0));
            else if (cmask == 1) LQ_NEAR_DENSE(1);
            else LQ_NEAR_DENSE(-1);
#undef LQ_NEAR_DENSE
        }
        launches++;
        if (nsplit_out) *nsplit_out = sa.nsplit;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }

    // S = lqr(sample, 0).S per query for state-dependent providers (planner.py:348), into parity `par`
    int launch_metric(const T* q, int qstride, int W, const int* d_count, int par, cudaStream_t st) {
        if (metric_diag || Lqr::CONST_METRIC || whiten_active()) return LQ_OK;
        if (W < 1) return LQ_OK;
        if (Lqr::COOP) {
            constexpr int ML = 8;
            metric_coop_kernel<T, Sys, Lqr, ML><<<(unsigned)(((long long)W * ML + 31) / 32), 32,
                                                  (32 / ML) * Lqr::SLOT_SCALARS * sizeof(T), st>>>(md, q, qstride, W, d_count, sp_of(par));
        } else {
            metric_kernel<T, Sys, Lqr><<<(W + 63) / 64, 64, 0, st>>>(md.sys, md.lqr, md.dt, q, qstride, W, d_count, sp_of(par));
        }
        launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }

    int launch_reduce(int par, int slot_lo, int nslots, int W, const int* d_count, cudaStream_t st) {
        if (nslots <= 1 || W < 1) return LQ_OK;
        reduce_partials_kernel<T><<<(W + 31) / 32, 256, 0, st>>>(pcost_of(par), pidx_of(par), Wc, slot_lo, nslots, W, d_count);
        launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }

    int scan_target() const { return 148 * (near_ctas_per_sm > 0 ? near_ctas_per_sm : (N <= 6 ? 16 : 8)); }
    // the old-tree scan shares the GPU with the rollout chain of the wave before: more, shorter CTAs let the chain's
    // (higher-priority) CTAs in sooner (measured on C4: 32 per SM 9.06 ms per plan, 16: 9.26, 8: 9.47)
    // pipelined waves keep two scans in flight, which fills the machine by itself: one resident set of CTAs per scan,
    // longer node chunks per CTA (measured on whole C4 plans, profiles/README.md: 8 per SM 6.57 ms, 16: 6.73, 4: 6.91)
    int pipelined_scan_target() const { return 148 * (near_ctas_per_sm > 0 ? near_ctas_per_sm : 8); }
    int old_scan_target() const { return near_ctas_per_sm > 0 ? 148 * near_ctas_per_sm : 2 * scan_target(); }

    int nearest(const void* d_q, int W, int use_ignore, int* d_idx, void* d_cost, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (W < 1 || W > Wc) LQ_FAIL(LQ_ERR_ARG, "W=%d outside 1..wave_capacity=%d", W, Wc);
        const T* q = (const T*)d_q;
        int qstride = W;
        if (int r = grid_sync(st)) return r;
        if (int r = whiten_sync(st)) return r;
        if (!metric_diag && !whiten_active()) {     // dense search reads S in the wave layout: restride the queries
            for (int d = 0; d < N; d++)
                LQ_CUDA(cudaMemcpyAsync(w_samples.p + (size_t)d * Wc, q + (size_t)d * W, W * sizeof(T), cudaMemcpyDeviceToDevice, st));
            q = w_samples.p; qstride = Wc;
            if (int r = launch_metric(q, qstride, W, nullptr, 0, st)) return r;
        }
        int nsplit = 1;
        // a full wave of queries fills the machine with one resident set of CTAs (ncu: 474 us at 8 per SM against 547 us
        // at 16 for 8192 queries x 1e5 nodes); small batches need the finer split
        if (int r = launch_scan(q, qstride, W, nullptr, nullptr, tv.size, size_ub, tv.n_ignored2[ign_copy], use_ignore, 0, 0, false, MAX_SPLIT,
                                W >= 4096 ? pipelined_scan_target() : scan_target(), st, &nsplit, ign_copy)) return r;
        nearest_finalize_kernel<T><<<(W + 127) / 128, 128, 0, st>>>(pcost_of(0), pidx_of(0), Wc, nsplit, W, d_idx, (T*)d_cost);
        launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }

    // ---- K2 -------------------------------------------------------------------
    // Rollouts have wildly different trip counts: small CTAs spread a wave over all 148 SMs, larger ones
    // amortise the shared-memory staging of the obstacle tables once every SM already holds several.
    template <class Kern>
    int steer_smem(Kern kern, size_t bytes) {
        if (bytes > 48 * 1024) LQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        return LQ_OK;
    }
    StagePlan stage_for_launch() const {
        StagePlan sp = stage_plan;
        if (!stage_tables_on) { memset(&sp, 0, sizeof sp); }
        return sp;
    }
    template <int LANES>
    int launch_steer_l(const int* ids, const T* targets, int tstride, int W, T* xs, T* us, int* len, cudaStream_t st) {
        const long long threads = (long long)W * LANES;
        const StagePlan sp = stage_for_launch();
        if (threads >= 148LL * 8 * 128 && !(Lqr::COOP && LANES > 1)) {
            const size_t sm = steer_coop_bytes<T, Lqr, LANES, 128>() + sp.bytes;
            auto kern = steer_kernel<T, Sys, Lqr, LANES, 128>;
            if (int r = steer_smem(kern, sm)) return r;
            kern<<<(unsigned)((threads + 127) / 128), 128, sm, st>>>(md, sp, t_state.p, t_K.p, cap, ids, targets, tstride, W, xs, us, len);
        } else {
            const size_t sm = steer_coop_bytes<T, Lqr, LANES, 32>() + sp.bytes;
            auto kern = steer_kernel<T, Sys, Lqr, LANES, 32>;
            if (int r = steer_smem(kern, sm)) return r;
            kern<<<(unsigned)((threads + 31) / 32), 32, sm, st>>>(md, sp, t_state.p, t_K.p, cap, ids, targets, tstride, W, xs, us, len);
        }
        launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }
    int launch_steer(const int* ids, const T* targets, int tstride, int W, T* xs, T* us, int* len, cudaStream_t st) {
        if (steer_lanes >= 32) return launch_steer_l<32>(ids, targets, tstride, W, xs, us, len, st);
        if (steer_lanes >= 8) return launch_steer_l<8>(ids, targets, tstride, W, xs, us, len, st);
        return launch_steer_l<1>(ids, targets, tstride, W, xs, us, len, st);
    }

    int steer(const int* d_ids, const void* d_t, int W, void* d_x, void* d_u, int* d_len, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (W < 1 || W > Wc) LQ_FAIL(LQ_ERR_ARG, "W=%d outside 1..wave_capacity=%d", W, Wc);
        T* xs = d_x ? (T*)d_x : w_xs.p;
        T* us = d_u ? (T*)d_u : w_us.p;
        if (d_x) LQ_CUDA(cudaMemsetAsync(xs, 0, (size_t)W * H * N * sizeof(T), st));
        if (d_u) LQ_CUDA(cudaMemsetAsync(us, 0, (size_t)W * H * M * sizeof(T), st));
        return launch_steer(d_ids, (const T*)d_t, W, W, xs, us, d_len ? d_len : w_len.p, st);
    }

    // The wave's rollout kernel: nearest selection from the partial slots [slot_lo, slot_lo + nslots) of parity `par`,
    // rollouts of the first ring[slot].count samples, results into `dst`.
    // copy: which copy of the ignore marks the wave searches; d_vis: the tree size it searches (device scalar);
    // xpar: which rollout scratch / completion counter it uses (two waves may be in flight with pipelined waves)
    template <int LANES>
    int launch_steer_x_l(int W, int slot, int par, int slot_lo, int nslots, int use_ignore, const ExDst& dst, int flags,
                         bool use_stop_word, cudaStream_t st, int copy, const int* d_vis, int xpar) {
        const long long threads = (long long)std::max(1, W) * LANES;
        const StagePlan sp = stage_for_launch();
        PartialsIn<T> pin{pcost_of(par) + (size_t)slot_lo * Wc, pidx_of(par) + (size_t)slot_lo * Wc, Wc, nslots};
        NearMetric<T, N> nm;
        nm.dmet = dmet; nm.dmet.angle_mask = md.angle_mask; nm.Spacked = sp_of(par); nm.diag = metric_diag ? 1 : 0;
        nm.zstate = nullptr; nm.U = nullptr;
        if (!metric_diag && whiten_active()) { nm.dmet = dmet_w; nm.diag = 2; nm.zstate = t_z.p; nm.U = w_U.p; }
        const int* stopw = use_stop_word ? d_stop : nullptr;
        int block = steer_block;
        if (block != 32 && block != 128) block = (threads >= 148LL * 2 * 128 && !(Lqr::COOP && LANES > 1)) ? 128 : 32;
        if (Lqr::COOP && LANES > 1) block = 32;
#define LQ_STEER_X(BB)                                                                                         \
    do {                                                                                                       \
        const size_t sm = steer_coop_bytes<T, Lqr, LANES, BB>() + sp.bytes;                                    \
        auto kern = steer_x_kernel<T, Sys, Lqr, LANES, BB>;                                                    \
        if (int r = steer_smem(kern, sm)) return r;                                                            \
        LQ_CUDA(launch_maybe_pdl(kern, (unsigned)((threads + BB - 1) / BB), BB, sm, st, pdl_on != 0,          \
            md, sp, t_state.p, t_K.p, cap, (const uint8_t*)tv.ignore[copy], d_vis, (const int*)tv.n_ignored2[copy], w_ctl.p, slot, pin, nm, \
            samples_of(par), Wc, tries_of(par), use_ignore, force_rescan, xs_of(xpar), us_of(xpar), dst, Wc, flags, stopw, xpar)); \
    } while (0)
        if (block == 128) LQ_STEER_X(128); else LQ_STEER_X(32);
#undef LQ_STEER_X
        launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }
    int launch_steer_x(int W, int slot, int par, int slot_lo, int nslots, int use_ignore, const ExDst& dst, int flags,
                       bool use_stop_word, cudaStream_t st, int copy = 0, const int* d_vis = nullptr, int xpar = 0) {
        if (!d_vis) d_vis = tv.size;
        if (steer_lanes >= 32) return launch_steer_x_l<32>(W, slot, par, slot_lo, nslots, use_ignore, dst, flags, use_stop_word, st, copy, d_vis, xpar);
        if (steer_lanes >= 8) return launch_steer_x_l<8>(W, slot, par, slot_lo, nslots, use_ignore, dst, flags, use_stop_word, st, copy, d_vis, xpar);
        return launch_steer_x_l<1>(W, slot, par, slot_lo, nslots, use_ignore, dst, flags, use_stop_word, st, copy, d_vis, xpar);
    }

    int launch_sample(uint64_t seed, uint32_t wave, int slot, int par, int count_ub, cudaStream_t st,
                      int first_known = 0, int count_known = -1) {
        constexpr int TL = 16;                       // >= tries_capacity
        const long long threads = (long long)std::max(1, count_ub) * TL;
        const WaveSlot* ws = &w_ctl.p->ring[slot];
        if (threads >= 148LL * 8 * 128) {
            sample_kernel<T, Sys, Lqr, TL, 128><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(
                md, sp, seed, wave, ws, samples_of(par), Wc, tries_of(par), first_known, count_known);
        } else {
            sample_kernel<T, Sys, Lqr, TL, 32><<<(unsigned)((threads + 31) / 32), 32, 0, st>>>(
                md, sp, seed, wave, ws, samples_of(par), Wc, tries_of(par), first_known, count_known);
        }
        launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }

    // ---- exchange destinations ------------------------------------------------
    // every rank's receive buffer: [PEER_PAR parities][world segments][seg_sz], then [PEER_PAR][world] arrival flags.
    // A wave uses parity seq % peer_mod: 2 suffice when a rank's waves are strictly sequential, pipelined waves (two in
    // flight per rank) need 4 so that a peer still committing wave k-2 is never overwritten by wave k's results.
    static constexpr int PEER_PAR = 4;
    int peer_mod = 2;
    size_t peer_flags_offset() const { return (size_t)PEER_PAR * peer_world * seg_sz; }
    int peer_parity() const { return (int)(peer_seq % (unsigned)peer_mod); }
    const char* peer_segs() const { return peer_buf[peer_rank] + (size_t)peer_parity() * peer_world * seg_sz; }
    const unsigned* peer_arrive() const { return (const unsigned*)(peer_buf[peer_rank] + peer_flags_offset()) + peer_parity() * peer_world; }
    // where this wave's rollout kernel sends its results (advances the wave sequence number in peer mode)
    ExDst next_dst(int par) {
        ExDst d;
        memset(&d, 0, sizeof d);
        if (peer_world <= 1) { d.n = 1; d.seg[0] = seg_of(par); return d; }
        peer_seq++;
        const int pp = peer_parity();
        d.n = peer_world; d.seq = peer_seq;
        for (int p = 0; p < peer_world; p++) {
            d.seg[p] = peer_buf[p] + ((size_t)pp * peer_world + peer_rank) * seg_sz;
            d.flag[p] = (unsigned*)(peer_buf[p] + peer_flags_offset()) + pp * peer_world + peer_rank;
        }
        return d;
    }
    int set_peers(int world, int rank, void* const* bufs, uint64_t bytes) override {
        if (world < 0 || world > MAX_PEERS || (world > 0 && (rank < 0 || rank >= world))) LQ_FAIL(LQ_ERR_ARG, "bad world/rank");
        if (world > 1) {
            const uint64_t need = (uint64_t)PEER_PAR * world * seg_sz + (uint64_t)PEER_PAR * world * sizeof(unsigned);
            if (!bufs || bytes < need) LQ_FAIL(LQ_ERR_ARG, "peer buffers must hold %llu bytes", (unsigned long long)need);
            for (int p = 0; p < world; p++) { if (!bufs[p]) LQ_FAIL(LQ_ERR_ARG, "null peer buffer"); peer_buf[p] = (char*)bufs[p]; }
        }
        peer_world = world > 1 ? world : 0; peer_rank = rank; peer_seq = 0;
        return LQ_OK;
    }

    // ---- one wave, stepwise driver (the host sizes and launches every wave) ------
    int expand_from_samples(int count, int use_ignore, int flags, cudaStream_t st) {
        const int slot = 0, par = cur_parity;
        const int ps = q_tail % RING;
        if (prof_on) cudaEventRecord(pev[ps][1], st);
        if (int r = grid_sync(st)) return r;
        if (int r = whiten_sync(st)) return r;
        if (int r = launch_metric(samples_of(par), Wc, count, nullptr, par, st)) return r;
        int nsplit = 1;
        // stepwise: one scan over the whole tree (n_old = 0), written to the "delta" slots
        if (int r = launch_scan(samples_of(par), Wc, count, nullptr, nullptr, tv.size, size_ub, nullptr, use_ignore, par,
                                MAX_SPLIT_OLD, false, MAX_SPLIT_NEW, scan_target(), st, &nsplit)) return r;
        if (prof_on) { cudaEventRecord(pev[ps][2], st); cudaEventRecord(pev[ps][3], st); }
        // stepwise: rollouts and commit of a wave run back to back on one stream, so the engine's own exchange segment
        // needs no double buffering (and the NCCL path's zero-copy send view keeps one fixed address)
        const ExDst dst = next_dst(0);
        if (int r = launch_steer_x(count, slot, par, MAX_SPLIT_OLD, nsplit, use_ignore, dst, flags, false, st)) return r;
        if (prof_on) { cudaEventRecord(pev[ps][4], st); prof_open[ps] = true; prof_W[ps] = count; }
        last_stream = st;
        return LQ_OK;
    }

    int wave_expand(uint64_t seed, uint32_t wave, int first, int count, const double* h_samples, int use_ignore,
                    int flags, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (count < 0 || count > Wc) LQ_FAIL(LQ_ERR_ARG, "count=%d outside 0..wave_capacity=%d", count, Wc);
        if (!have_goal) LQ_FAIL(LQ_ERR_STATE, "lq_set_goal has not been called");
        if (q_tail - q_head >= (unsigned)RING) LQ_FAIL(LQ_ERR_STATE, "%d committed waves are unread: call lq_wave_stats_get", RING);
        const int ps = q_tail % RING;
        if (prof_on) cudaEventRecord(pev[ps][0], st);
        cur_parity = 0;
        if (!h_samples && count > 0 && prefetch_on) { cur_parity = sample_parity; }
        ctl_set_kernel<<<1, 1, 0, st>>>(w_ctl.p, 0, count, first, count, 0); launches++;
        if (count > 0 && h_samples) {
            std::vector<T> soa((size_t)N * count);
            for (int j = 0; j < count; j++)
                for (int d = 0; d < N; d++) soa[(size_t)d * count + j] = (T)h_samples[(size_t)j * N + d];
            for (int d = 0; d < N; d++)
                LQ_CUDA(cudaMemcpyAsync(samples_of(0) + (size_t)d * Wc, soa.data() + (size_t)d * count, count * sizeof(T),
                                        cudaMemcpyHostToDevice, st));
            LQ_CUDA(cudaMemsetAsync(tries_of(0), 0, count * sizeof(int), st));
            LQ_CUDA(cudaStreamSynchronize(st));
        } else if (count > 0) {
            if (!have_sampling) LQ_FAIL(LQ_ERR_STATE, "lq_set_sampling has not been called");
            if (prefetch_on) {
                // The sampler reads nothing of the tree (planner.py:201-211), and the host enqueues wave k+1
                // while wave k is still running: issue it on the side stream into the other buffer, so it
                // overlaps wave k's scan / rollouts instead of heading wave k+1's dependent chain.  The slice
                // is passed through a side-stream copy of the control slot (slot 1).
                if (int r = prefetch_setup()) return r;
                const int par = sample_parity; sample_parity ^= 1;
                if (consumed_valid[par]) LQ_CUDA(cudaStreamWaitEvent(aux, ev_consumed[par], 0));   // its last reader (two waves ago)
                ctl_set_kernel<<<1, 1, 0, aux>>>(w_ctl.p, 2 + par, count, first, count, 0); launches++;
                if (int r = launch_sample(seed, wave, 2 + par, par, count, aux)) return r;
                LQ_CUDA(cudaEventRecord(ev_sampled[par], aux));
                LQ_CUDA(cudaStreamWaitEvent(st, ev_sampled[par], 0));
                const int r = expand_from_samples(count, use_ignore, flags, st);
                LQ_CUDA(cudaEventRecord(ev_consumed[par], st));
                consumed_valid[par] = true;
                return r;
            }
            if (int r = launch_sample(seed, wave, 0, 0, count, st)) return r;
        }
        return expand_from_samples(count, use_ignore, flags, st);
    }

    int wave_expand_candidates(int count, int tries, const double* h_cand, int n_cand, int use_ignore,
                               int* h_consumed, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (count < 1 || count > Wc) LQ_FAIL(LQ_ERR_ARG, "count=%d outside 1..wave_capacity=%d", count, Wc);
        if (tries < 1 || n_cand < (long long)count * tries) LQ_FAIL(LQ_ERR_ARG, "need count*tries candidates");
        if (!have_goal) LQ_FAIL(LQ_ERR_STATE, "lq_set_goal has not been called");
        if (w_cand.n < (size_t)N * n_cand) { LQ_CUDA(w_cand.alloc((size_t)N * n_cand)); LQ_CUDA(w_candok.alloc(n_cand)); }
        if (q_tail - q_head >= (unsigned)RING) LQ_FAIL(LQ_ERR_STATE, "%d committed waves are unread: call lq_wave_stats_get", RING);
        if (prof_on) cudaEventRecord(pev[q_tail % RING][0], st);
        cur_parity = 0;
        ctl_set_kernel<<<1, 1, 0, st>>>(w_ctl.p, 0, count, 0, count, 0); launches++;
        std::vector<T> soa((size_t)N * n_cand);
        for (int j = 0; j < n_cand; j++)
            for (int d = 0; d < N; d++) soa[(size_t)d * n_cand + j] = (T)h_cand[(size_t)j * N + d];
        LQ_CUDA(cudaMemcpyAsync(w_cand.p, soa.data(), soa.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        candidates_feasible_kernel<T, Sys, Lqr><<<(n_cand + 127) / 128, 128, 0, st>>>(md, w_cand.p, n_cand, n_cand, w_candok.p);
        launches++;
        candidates_select_kernel<T, N><<<1, 256, 0, st>>>(w_cand.p, n_cand, w_candok.p, n_cand, count, tries,
                                                          samples_of(0), Wc, w_pick.p, w_misc.p);
        launches++;
        LQ_CUDA(cudaMemsetAsync(tries_of(0), 0, count * sizeof(int), st));
        LQ_CUDA(cudaMemcpyAsync(h_misc, w_misc.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (int r = expand_from_samples(count, use_ignore, 0, st)) return r;
        LQ_CUDA(cudaStreamSynchronize(st));
        if (h_consumed) *h_consumed = h_misc[0];
        return LQ_OK;
    }

    int wave_records(void** d, uint64_t* bytes) override {
        if (d) *d = seg_of(0);
        if (bytes) *bytes = seg_sz;
        return LQ_OK;
    }

    int launch_commit(int wave_k, const char* segs, int n_seg, int my_seg, int slots_ub, const WavePolicy& pol,
                      const unsigned* arrive, unsigned seq, int ring_slot, unsigned out_seq, cudaStream_t st,
                      int copy = 0, int lag2 = 0, int xpar = 0, bool pdl = true) {
        ign_copy = copy;
        const int ctas = std::max(1, (slots_ub + COMMIT_SLOTS - 1) / COMMIT_SLOTS);
        LQ_CUDA(launch_maybe_pdl(commit_x_kernel<T>, (unsigned)ctas, COMMIT_BLOCK, 0, st, pdl && pdl_on != 0 && pol.device_schedule != 0,
                                 tv, w_ctl.p, wave_k, segs, seg_sz, n_seg, my_seg, (const T*)xs_of(xpar), (const T*)us_of(xpar),
                                 Wc, N, M, H, (const T*)w_goal.p, (const T*)(w_goal.p + LQ_MAX_N), pol, d_ring + ring_slot, out_seq,
                                 arrive, seq, w_hor.p, md.hfactor, md.hlo, md.hhi, copy, lag2));
        launches++;
        LQ_CUDA(cudaGetLastError());
        return LQ_OK;
    }

    int wave_commit(const void* d_records, int n_seg, int my_seg, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (n_seg < 1 || n_seg > 255) LQ_FAIL(LQ_ERR_ARG, "n_segments must be in 1..255");
        if (my_seg < 0 || my_seg >= n_seg) LQ_FAIL(LQ_ERR_ARG, "my_segment must be in 0..n_segments-1");
        last_my_seg = my_seg;
        const char* segs = d_records ? (const char*)d_records : seg_of(0);
        const unsigned* arrive = nullptr;
        if (!d_records && peer_world > 1) {             // fused exchange: my own receive buffer, current parity
            if (n_seg != peer_world || my_seg != peer_rank) LQ_FAIL(LQ_ERR_ARG, "peer mode commits world segments as rank");
            segs = peer_segs();
            arrive = peer_arrive();
        }
        if (q_tail - q_head >= (unsigned)RING) LQ_FAIL(LQ_ERR_STATE, "%d committed waves are unread: call lq_wave_stats_get", RING);
        const int ps = q_tail % RING;
        WavePolicy pol;
        memset(&pol, 0, sizeof pol);
        pol.world = 1; pol.device_schedule = 0;
        if (prof_on) cudaEventRecord(pev[ps][5], st);
        ++stat_seq;
        if (int r = launch_commit(0, segs, n_seg, my_seg, n_seg * Wc, pol, arrive, peer_seq, ps, stat_seq, st)) return r;
        if (prof_on) cudaEventRecord(pev[ps][6], st);
        q_seq[ps] = stat_seq; q_stream[ps] = st;
        q_add[ps] = n_seg * Wc;
        q_tail++;
        size_ub = std::min(cap, size_ub + n_seg * Wc);
        last_stream = st;
        return LQ_OK;
    }

    // Block until the commit that carries sequence number `seq` has published ring slot `ps`.
    int wait_stats(int ps, unsigned seq, cudaStream_t st) {
        volatile unsigned* s = &h_ring[ps].seq;
        unsigned spins = 0;
        while ((int)(*s - seq) < 0) {
            if ((++spins & 0x3ff) == 0) {
                const cudaError_t e = cudaStreamQuery(st);
                if (e == cudaSuccess) {
                    if ((int)(*s - seq) >= 0) break;
                    LQ_FAIL(LQ_ERR_STATE, "the stream drained but the wave counters never arrived");
                }
                if (e != cudaErrorNotReady) LQ_CUDA(e);
            }
        }
        if (h_ring[ps].v[10]) LQ_FAIL(LQ_ERR_STATE, "a peer GPU never delivered its wave records (exchange timed out)");
        return LQ_OK;
    }

    void account_profile(int ps, int local_count, int size_before) {
        if (!prof_on || !prof_open[ps]) return;
        float ms = 0;
        auto el = [&](int a, int b) { ms = 0; cudaEventElapsedTime(&ms, pev[ps][a], pev[ps][b]); return (double)ms; };
        prof_ms[0] += el(0, 1); prof_ms[1] += el(prof_scan_ev[ps] ? prof_scan_ev[ps] : 1, 2); prof_ms[2] += el(3, 4) - 0.0; prof_ms[3] += el(4, 5); prof_ms[4] += el(5, 6);
        prof_work[0] += (double)local_count * size_before;
        prof_work[1] += (double)local_count * size_before * N * sizeof(T);
        prof_work[2] += local_count; prof_work[3] += 1;
        prof_open[ps] = false;
    }

    int wave_stats(lq_wave_stats* out) override {
        LQ_CUDA(cudaSetDevice(device));
        if (q_head != q_tail) {                       // oldest committed, unread wave
            const int ps = q_head % RING;
            const int size_before = size_ub_exact;
            if (int r = wait_stats(ps, q_seq[ps], q_stream[ps])) return r;
            if (prof_on && prof_open[ps]) {
                LQ_CUDA(cudaEventSynchronize(pev[ps][6]));
                // stepwise events: 0 start | 1 sampled | 2 scanned (= 3) | 4 rolled out | 5 (= 4) | 6 committed
                float ms = 0;
                cudaEventElapsedTime(&ms, pev[ps][0], pev[ps][1]); prof_ms[0] += ms;
                cudaEventElapsedTime(&ms, pev[ps][1], pev[ps][2]); prof_ms[2] += ms;
                cudaEventElapsedTime(&ms, pev[ps][3], pev[ps][4]); prof_ms[3] += ms;
                cudaEventElapsedTime(&ms, pev[ps][5], pev[ps][6]); prof_ms[4] += ms;
                prof_work[0] += (double)prof_W[ps] * size_before;
                prof_work[1] += (double)prof_W[ps] * size_before * N * sizeof(T);
                prof_work[2] += prof_W[ps]; prof_work[3] += 1;
                prof_open[ps] = false;
            }
            memcpy(&last_stats, h_ring[ps].v, sizeof last_stats);
            q_head++;
            size_ub_exact = last_stats.tree_size;
            int ub = last_stats.tree_size;                // exact size + what the waves still in flight may add
            for (unsigned q = q_head; q != q_tail; q++) ub += q_add[q % RING];
            size_ub = std::min(cap, ub);
        }
        if (out) *out = last_stats;
        return LQ_OK;
    }

    // ---- the device-scheduled driver (lq_plan_run) -------------------------------------------------------
    // The whole expand loop of one plan (planner.py:233-311) as a pipeline the host only feeds: wave sizes, slices,
    // the old/new node split of the scan and the stop decision live in the device control block, so waves are
    // enqueued DEPTH ahead of the counters the host has seen.  Per wave k:
    //   scan stream   wait commit k-2 | sampler k | [metric k] | old-tree scan k over [0, N_{k-1})
    //   chain stream  wait scan k     | delta scan k over [N_{k-1}, N_k) | rollouts + exchange k | commit k
    // so wave k+1's scan of the old tree runs while wave k's rollouts, exchange and commit are in flight, and the
    // dependent chain of a wave is delta scan -> rollouts -> commit.  The tree is the one the stepwise driver grows
    // with the same wave sizes (lexicographic merge of the partial minima, per-query check of the ignore marks).
    int ensure_streams() {
        if (s_chain) return LQ_OK;
        int lo = 0, hi = 0;
        LQ_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority (numerically greatest)
        LQ_CUDA(cudaStreamCreateWithPriority(&s_chain, cudaStreamNonBlocking, hi));
        LQ_CUDA(cudaStreamCreateWithPriority(&s_scan, cudaStreamNonBlocking, lo));
        LQ_CUDA(cudaStreamCreateWithPriority(&s_chain2, cudaStreamNonBlocking, hi));
        // the sampler that runs ahead of a pipelined wave's chain: short, and the wave's scan waits for it, so it must not
        // queue behind the long rollout kernels of the other chain (measured on C5, where a low-priority sampler starved)
        LQ_CUDA(cudaStreamCreateWithPriority(&s_samp, cudaStreamNonBlocking, hi));
        for (auto& e : ev_commit) LQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ev_scan) LQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ev_steer) LQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ev_samp) LQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        LQ_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        LQ_CUDA(cudaEventCreate(&ev_begin));
        LQ_CUDA(cudaEventCreate(&ev_end));
        return LQ_OK;
    }

    int plan_run(const lq_plan_args* a, lq_plan_summary* out, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (!a || !out) LQ_FAIL(LQ_ERR_ARG, "null argument");
        memset(out, 0, sizeof *out);
        if (!have_goal) LQ_FAIL(LQ_ERR_STATE, "lq_set_goal has not been called");
        if (!have_sampling) LQ_FAIL(LQ_ERR_STATE, "lq_set_sampling has not been called");
        if (q_head != q_tail) LQ_FAIL(LQ_ERR_STATE, "stepwise waves are still unread");
        const int world = std::max(1, a->world), rank = a->rank;
        if (rank < 0 || rank >= world) LQ_FAIL(LQ_ERR_ARG, "bad world/rank");
        if (world > 1 && (peer_world != world || peer_rank != rank)) LQ_FAIL(LQ_ERR_STATE, "lq_plan_run on several GPUs needs lq_wave_set_peers first");
        if (world == 1 && peer_world > 1) LQ_FAIL(LQ_ERR_STATE, "peer buffers are set for %d ranks", peer_world);
        WavePolicy pol;
        memset(&pol, 0, sizeof pol);
        pol.fixed = a->wave_fixed; pol.wmin = a->wave_min; pol.wmax = a->wave_max; pol.growth = a->growth;
        pol.world = world; pol.rank = rank; pol.device_schedule = 1;
        const int wtop = pol.fixed > 0 ? pol.fixed : pol.wmax;
        if (wtop < 1 || (wtop + world - 1) / world > Wc) LQ_FAIL(LQ_ERR_ARG, "largest wave %d over %d ranks exceeds wave_capacity %d", wtop, world, Wc);
        if (pol.fixed <= 0 && (pol.wmin < 1 || pol.wmax < pol.wmin || !(pol.growth > 0))) LQ_FAIL(LQ_ERR_ARG, "bad wave policy");
        if (int r = ensure_streams()) return r;
        const int DEPTH = std::max(1, std::min(a->depth > 0 ? a->depth : 4, RING - 2));
        const bool spec = spec_scan != 0;
        const int use_ignore = a->use_ignore ? 1 : 0;
        const bool lag2 = a->lag == 2;
        if (a->lag != 0 && a->lag != 1 && a->lag != 2) LQ_FAIL(LQ_ERR_ARG, "lag must be 1 or 2");
        peer_mod = lag2 ? 4 : 2;
        if (peer_world > 1 && (peer_seq % 4u) != 0) peer_seq += 4u - (peer_seq % 4u);     // every rank alike: start on parity 0
        last_my_seg = rank;

        // the chain stream picks up after the caller's stream (reset / goal / sampling uploads)
        if (int r = grid_sync(st)) return r;
        if (int r = whiten_sync(st)) return r;
        LQ_CUDA(cudaEventRecord(ev_in, st));
        LQ_CUDA(cudaStreamWaitEvent(s_chain, ev_in, 0));
        *h_stop = 0;
        ctl_start_kernel<<<1, 1, 0, s_chain>>>(w_ctl.p, tv.size, pol); launches++;
        LQ_CUDA(cudaEventRecord(ev_begin, s_chain));
        LQ_CUDA(cudaStreamWaitEvent(s_scan, ev_begin, 0));         // waves 0 and 1 start from the initial control block
        LQ_CUDA(cudaStreamWaitEvent(s_chain2, ev_begin, 0));
        LQ_CUDA(cudaStreamWaitEvent(s_samp, ev_begin, 0));

        // host mirror of the schedule as upper bounds (they only shape grids): Nub[k] >= N_k, the tree size wave k
        // starts from, and Wub[k] >= W_k = policy(N_{k-1}) (W_0 = W_1 = policy(N_0))
        const int n0 = size_ub_exact;
        std::vector<int> Nub, Wub;
        std::vector<int> n_start;                    // exact N_k once known (for the profile's pair counts)
        int enq = 0, rd = 0;
        bool stop_enqueue = false;
        long long samples_total = 0;
        int exact_upto = 0;                          // n_start[k] = N_k is known for k <= exact_upto
        memset(&last_stats, 0, sizeof last_stats);
        last_stats.tree_size = n0; last_stats.best_node = -1; last_stats.best_cost = -1;
        std::vector<unsigned> seq_of_wave;

        struct NvtxRange { explicit NvtxRange(const char* n) { nvtxRangePushA(n); } ~NvtxRange() { nvtxRangePop(); } };
        NvtxRange plan_range("lq_plan_run");
        auto enqueue = [&](int k) -> int {
            NvtxRange wave_range(lag2 ? "lq wave (pipelined): sample|scan|fold|rollout+exchange|commit" : "lq wave: sample|old scan|fold || delta scan|rollout+exchange|commit");
            // bounds for wave k, re-derived from the newest exact size
            if ((int)Nub.size() <= k) { Nub.resize(k + 1); Wub.resize(k + 1); }
            for (int i = std::max(0, exact_upto - 2); i <= k; i++) {
                Nub[i] = i <= exact_upto ? n_start[i] : std::min(tv.limit, Nub[i - 1] + Wub[i - 1]);
                Wub[i] = policy_wave(pol, i < 2 ? n0 : Nub[i - 1]);
            }
            const int w = Wub[k];
            const int cub = (w + world - 1) / world;
            const int slot = k % CTL_RING, par = k & 1, ps = k % RING;
            const int n_old_ub = k == 0 ? n0 : Nub[k - 1];
            const int delta_ub = k == 0 ? 1 : std::max(1, std::min(Wub[k - 1], tv.limit));
            const int* d_count = &w_ctl.p->ring[slot].count;
            const int* d_nold = &w_ctl.p->ring[slot].n_old;
            if (lag2) {
                // Pipelined waves: wave k lives on chain stream k & 1 -- sampler, scan of the tree as of commit k-2
                // ([0, n_old): that is everything the wave may see, so there is no delta scan), fold of the partial minima,
                // rollouts + exchange -- and only its commit waits for commit k-1.  Wave k+1's whole front half runs on the
                // other stream meanwhile.
                cudaStream_t sc = (k & 1) ? s_chain2 : s_chain;
                const int copy = k & 1;
                // The sampler reads nothing of the tree (planner.py:201-211), only the wave's size.  Once the policy has
                // reached its top size (it is monotone in the tree size), and for waves 0 and 1, the host knows that size
                // itself: the sampler then runs on the side stream as soon as the sample buffer's last reader (the
                // rollouts two waves back) is done, off the wave's chain.  After a stop the slot says 0 samples and the
                // rollout kernel ignores what was drawn.
                const int n_known = n_start[exact_upto];
                const bool ahead = sample_ahead && (pol.fixed > 0 || k < 2 || policy_wave(pol, n_known) >= pol.wmax);
                if (ahead) {
                    const int wk = k < 2 ? policy_wave(pol, n0) : (pol.fixed > 0 ? pol.fixed : pol.wmax);
                    int first_k = 0, count_k = 0;
                    policy_slice(wk, rank, world, &first_k, &count_k);
                    if (k >= 2) LQ_CUDA(cudaStreamWaitEvent(s_samp, ev_steer[(k - 2) % EVR], 0));
                    if (prof_on) cudaEventRecord(pev[ps][0], s_samp);
                    if (int r = launch_sample(a->seed, a->wave_ctr0 + (uint32_t)k, slot, par, cub, s_samp, first_k, count_k)) return r;
                    if (prof_on) cudaEventRecord(pev[ps][1], s_samp);
                    LQ_CUDA(cudaEventRecord(ev_samp[k % EVR], s_samp));
                    LQ_CUDA(cudaStreamWaitEvent(sc, ev_samp[k % EVR], 0));
                } else {
                    if (prof_on) cudaEventRecord(pev[ps][0], sc);
                    if (int r = launch_sample(a->seed, a->wave_ctr0 + (uint32_t)k, slot, par, cub, sc)) return r;
                    if (prof_on) cudaEventRecord(pev[ps][1], sc);
                }
                if (prof_on) { cudaEventRecord(pev[ps][7], sc); prof_scan_ev[ps] = 7; }
                if (int r = launch_metric(samples_of(par), Wc, cub, d_count, par, sc)) return r;
                int ns = 0;
                if (int r = launch_scan(samples_of(par), Wc, cub, d_count, nullptr, d_nold, n_old_ub, nullptr, use_ignore, par,
                                        MAX_SPLIT_OLD, true, MAX_SPLIT_OLD, pipelined_scan_target(), sc, &ns, copy)) return r;
                if (int r = launch_reduce(par, MAX_SPLIT_OLD - ns, ns, cub, d_count, sc)) return r;
                if (prof_on) { cudaEventRecord(pev[ps][2], sc); cudaEventRecord(pev[ps][3], sc); cudaEventRecord(pev[ps][4], sc); }
                const ExDst dst = next_dst(par);
                if (int r = launch_steer_x(cub, slot, par, MAX_SPLIT_OLD - 1, 1, use_ignore, dst, 0, true, sc, copy, d_nold, par)) return r;
                LQ_CUDA(cudaEventRecord(ev_steer[k % EVR], sc));
                if (prof_on) cudaEventRecord(pev[ps][5], sc);
                if (k >= 1) LQ_CUDA(cudaStreamWaitEvent(sc, ev_commit[(k - 1) % EVR], 0));
                const char* segs = seg_of(par);
                const unsigned* arrive = nullptr;
                if (peer_world > 1) { segs = peer_segs(); arrive = peer_arrive(); }
                ++stat_seq;
                seq_of_wave.push_back(stat_seq);
                if (int r = launch_commit(k, segs, world, rank, w, pol, arrive, peer_seq, ps, stat_seq, sc, copy, 1, par, false)) return r;
                if (prof_on) { cudaEventRecord(pev[ps][6], sc); prof_open[ps] = true; }
                LQ_CUDA(cudaEventRecord(ev_commit[k % EVR], sc));
                return LQ_OK;
            }
            // ---- scan stream: sampler, [metric], old-tree scan
            if (k >= 2) LQ_CUDA(cudaStreamWaitEvent(s_scan, ev_commit[(k - 2) % EVR], 0));
            if (prof_on) cudaEventRecord(pev[ps][0], s_scan);
            if (int r = launch_sample(a->seed, a->wave_ctr0 + (uint32_t)k, slot, par, cub, s_scan)) return r;
            if (int r = launch_metric(samples_of(par), Wc, cub, d_count, par, s_scan)) return r;
            if (prof_on) { cudaEventRecord(pev[ps][1], s_scan); prof_scan_ev[ps] = 1; }
            int ns_old = 0, ns_new = 0;
            if (spec) {
                if (int r = launch_scan(samples_of(par), Wc, cub, d_count, nullptr, d_nold, n_old_ub, nullptr, use_ignore, par,
                                        MAX_SPLIT_OLD, true, MAX_SPLIT_OLD, old_scan_target(), s_scan, &ns_old)) return r;
            }
            if (spec) {
                // fold the old-tree partials into one slot here, off the rollout chain
                if (int r = launch_reduce(par, MAX_SPLIT_OLD - ns_old, ns_old, cub, d_count, s_scan)) return r;
                if (ns_old > 0) ns_old = 1;
            }
            if (prof_on) cudaEventRecord(pev[ps][2], s_scan);
            LQ_CUDA(cudaEventRecord(ev_scan[k % EVR], s_scan));
            // ---- chain stream: delta scan (the whole tree when not speculating), rollouts + exchange, commit
            LQ_CUDA(cudaStreamWaitEvent(s_chain, ev_scan[k % EVR], 0));
            if (prof_on) cudaEventRecord(pev[ps][3], s_chain);
            if (spec) {
                if (int r = launch_scan(samples_of(par), Wc, cub, d_count, d_nold, tv.size, delta_ub, nullptr, use_ignore, par,
                                        MAX_SPLIT_OLD, false, MAX_SPLIT_NEW, scan_target() / 2, s_chain, &ns_new)) return r;
            } else {
                if (int r = launch_scan(samples_of(par), Wc, cub, d_count, nullptr, tv.size, Nub[k], nullptr, use_ignore, par,
                                        MAX_SPLIT_OLD, false, MAX_SPLIT_NEW, scan_target(), s_chain, &ns_new)) return r;
            }
            if (prof_on) cudaEventRecord(pev[ps][4], s_chain);
            const ExDst dst = next_dst(par);
            if (int r = launch_steer_x(cub, slot, par, MAX_SPLIT_OLD - ns_old, ns_old + ns_new, use_ignore, dst, 0, true, s_chain)) return r;
            if (prof_on) cudaEventRecord(pev[ps][5], s_chain);
            const char* segs = seg_of(par);
            const unsigned* arrive = nullptr;
            if (peer_world > 1) { segs = peer_segs(); arrive = peer_arrive(); }
            ++stat_seq;
            seq_of_wave.push_back(stat_seq);
            if (int r = launch_commit(k, segs, world, rank, w, pol, arrive, peer_seq, ps, stat_seq, s_chain)) return r;
            if (prof_on) { cudaEventRecord(pev[ps][6], s_chain); prof_open[ps] = true; }
            LQ_CUDA(cudaEventRecord(ev_commit[k % EVR], s_chain));
            return LQ_OK;
        };

        // Every rank enqueues wave k only after it has seen the counters of wave k - DEPTH without a stop in them, and
        // the stop (rank 0's request, or a full tree) is published by the commits identically on every rank -- so all
        // ranks enqueue the same number of waves and nobody waits for a peer that has left.
        n_start.push_back(n0);
        int rc = LQ_OK;
        using clk = std::chrono::steady_clock;
        auto ms_since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
        double t_enq = 0, t_wait = 0, t_cb = 0;
        while (true) {
            auto t0 = clk::now();
            while (!stop_enqueue && enq - rd < DEPTH) {
                if ((rc = enqueue(enq)) != LQ_OK) break;
                enq++;
            }
            t_enq += ms_since(t0);
            if (rc != LQ_OK || rd == enq) break;
            const int ps = rd % RING;
            t0 = clk::now();
            NvtxRange wait_range("lq wait for wave counters");
            rc = wait_stats(ps, seq_of_wave[rd], (lag2 && (rd & 1)) ? s_chain2 : s_chain);
            t_wait += ms_since(t0);
            if (rc != LQ_OK) break;
            lq_wave_stats stw;
            memcpy(&stw, h_ring[ps].v, sizeof stw);
            if (prof_on && prof_open[ps]) {
                cudaEventSynchronize(pev[ps][6]);
                int first = 0, cnt = 0;
                policy_slice(stw.n_sampled, rank, world, &first, &cnt);
                account_profile(ps, cnt, n_start[rd]);
            }
            samples_total += stw.n_sampled;
            exact_upto = rd + 1;
            n_start.push_back(stw.tree_size);
            last_stats = stw;
            out->waves = rd + 1;
            if (stw.flags) stop_enqueue = true;                 // the device has stopped: later waves are no-ops
            if (a->on_wave) {
                t0 = clk::now();
                const int want = a->on_wave(a->user, &stw);     // planner.py:286-311 on the host's clock
                t_cb += ms_since(t0);
                if (want && !stop_enqueue) *(volatile int*)h_stop = want;   // reaches every rank through rank 0's segment header
            }
            rd++;
        }
        if (rc != LQ_OK) {                                       // leave nothing in flight behind an error
            cudaStreamSynchronize(s_scan); cudaStreamSynchronize(s_samp); cudaStreamSynchronize(s_chain2); cudaStreamSynchronize(s_chain);
            return rc;
        }
        if (lag2 && enq > 0) {                                   // the last commit may sit on the second chain stream
            LQ_CUDA(cudaStreamWaitEvent(s_chain, ev_commit[(enq - 1) % EVR], 0));
        }
        LQ_CUDA(cudaEventRecord(ev_end, s_chain));
        LQ_CUDA(cudaStreamSynchronize(s_scan));
        LQ_CUDA(cudaStreamSynchronize(s_samp));
        LQ_CUDA(cudaStreamSynchronize(s_chain2));
        LQ_CUDA(cudaStreamSynchronize(s_chain));
        LQ_CUDA(cudaStreamWaitEvent(st, ev_end, 0));            // later work on the caller's stream sees the finished tree
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        out->device_ms = ms;
        out->host_enqueue_ms = t_enq; out->host_wait_ms = t_wait; out->host_callback_ms = t_cb;
        out->waves_enqueued = enq;
        out->samples = samples_total;
        out->last = last_stats;
        out->stop_code = last_stats.flags;
        size_ub = size_ub_exact = last_stats.tree_size;
        last_stream = st;
        return LQ_OK;
    }

    int stop_word(int** outp) override { *outp = h_stop; return LQ_OK; }

    // ---- close-out ------------------------------------------------------------
    int closest(const double* xg, const uint8_t* colmask, int* id, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        // Sguide = lqr(xguide, 0).S with masked columns zeroed (planner.py:313-316)
        FullMetric<T, N> fm;
        double S[N * N], K[M * N], u0[M] = {0};
        if (int r = eval_lqr(1, xg, u0, S, K)) return r;
        for (int i = 0; i < N; i++) {
            fm.x[i] = (T)xg[i];
            for (int j = 0; j < N; j++) fm.S[i * N + j] = (colmask && colmask[j]) ? T(0) : (T)S[i * N + j];
        }
        fm.angle_mask = md.angle_mask;
        closest_kernel<T, N><<<CLOSEST_CTAS, CLOSEST_BLOCK, 0, st>>>(t_state.p, cap, tv.size, fm, w_misc.p + 1, pcost_of(0), pidx_of(0), (unsigned*)(w_misc.p + 4)); launches++;
        LQ_CUDA(cudaGetLastError());
        LQ_CUDA(cudaMemcpyAsync(h_misc + 1, w_misc.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        LQ_CUDA(cudaStreamSynchronize(st));
        *id = h_misc[1];
        return LQ_OK;
    }

    int steer_force(int node, const double* xtar, int max_steps, double* hx, double* hu, int* len, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (max_steps < 1) LQ_FAIL(LQ_ERR_ARG, "max_steps must be >= 1");
        int sz = 0;
        if (int r = tree_size(&sz)) return r;
        if (node < 0 || node >= sz) LQ_FAIL(LQ_ERR_ARG, "The given ID, %d, doesn't exist.", node);
        DevBuf<T> xt, xs, us;
        std::vector<T> xth(N);
        for (int d = 0; d < N; d++) xth[d] = (T)xtar[d];
        if (int r = upload(xt, xth)) return r;
        LQ_CUDA(xs.alloc((size_t)max_steps * N)); LQ_CUDA(us.alloc((size_t)max_steps * M));
        steer_force_kernel<T, Sys, Lqr><<<1, 32, 0, st>>>(md, t_state.p, t_K.p, cap, node, xt.p, max_steps, xs.p, us.p, w_misc.p + 2);
        launches++;
        LQ_CUDA(cudaGetLastError());
        LQ_CUDA(cudaStreamSynchronize(st));
        int L = 0;
        LQ_CUDA(cudaMemcpy(&L, w_misc.p + 2, sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<T> bx((size_t)std::max(1, L) * N), bu((size_t)std::max(1, L) * M);
        if (L > 0) {
            LQ_CUDA(cudaMemcpy(bx.data(), xs.p, (size_t)L * N * sizeof(T), cudaMemcpyDeviceToHost));
            LQ_CUDA(cudaMemcpy(bu.data(), us.p, (size_t)L * M * sizeof(T), cudaMemcpyDeviceToHost));
            for (size_t k = 0; k < (size_t)L * N; k++) hx[k] = (double)bx[k];
            for (size_t k = 0; k < (size_t)L * M; k++) hu[k] = (double)bu[k];
        }
        *len = L;
        return LQ_OK;
    }

    int append_node(int parent, const double* state, const double* hx, const double* hu, int len, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        int sz = 0;
        if (int r = tree_size(&sz)) return r;
        if (parent < 0 || parent >= sz) LQ_FAIL(LQ_ERR_ARG, "The given parent ID, %d, doesn't exist.", parent);   // tree.py:83-84
        if (sz >= cap) LQ_FAIL(LQ_ERR_STATE, "tree storage is full (%d rows)", cap);
        // edges longer than the horizon (force-arrive) keep only their length on the device; the host
        // planner holds the full sequence (planner.py:299-303 extends its own lists)
        const int keep = std::min(len, H);
        std::vector<T> ex((size_t)H * N, T(0)), eu((size_t)H * M, T(0)), xs(N);
        for (int k = 0; k < keep * N; k++) ex[k] = (T)hx[k];
        for (int k = 0; k < keep * M; k++) eu[k] = (T)hu[k];
        for (int d = 0; d < N; d++) xs[d] = (T)state[d];
        for (int d = 0; d < N; d++) LQ_CUDA(cudaMemcpy(t_state.p + (size_t)d * cap + sz, &xs[d], sizeof(T), cudaMemcpyHostToDevice));
        LQ_CUDA(cudaMemcpy(t_ex.p + (size_t)sz * H * N, ex.data(), ex.size() * sizeof(T), cudaMemcpyHostToDevice));
        LQ_CUDA(cudaMemcpy(t_eu.p + (size_t)sz * H * M, eu.data(), eu.size() * sizeof(T), cudaMemcpyHostToDevice));
        int pc = 0;
        LQ_CUDA(cudaMemcpy(&pc, t_cost.p + parent, sizeof(int), cudaMemcpyDeviceToHost));
        const int c = pc + len;
        LQ_CUDA(cudaMemcpy(t_parent.p + sz, &parent, sizeof(int), cudaMemcpyHostToDevice));
        LQ_CUDA(cudaMemcpy(t_cost.p + sz, &c, sizeof(int), cudaMemcpyHostToDevice));
        LQ_CUDA(cudaMemcpy(t_elen.p + sz, &keep, sizeof(int), cudaMemcpyHostToDevice));
        // every rank appends this node itself (the edge arrives from the host): owner 255 = "held by all"
        const uint8_t own = 255;
        LQ_CUDA(cudaMemcpy(t_owner.p + sz, &own, 1, cudaMemcpyHostToDevice));
        // the reference stores lqr=None for this node (planner.py:299); K = lqr(state, 0) keeps it steerable
        init_gain_kernel<T, Sys, Lqr><<<1, 32, 0, st>>>(md, t_state.p + sz, t_K.p + sz, cap, 1); launches++;
        if (whiten_on && !z_dirty) { if (int r = whiten_range(sz, nullptr, 1, st)) return r; }
        if (grid_active() && !grid_dirty) { grid_insert_range_kernel<T><<<1, 32, 0, st>>>(tv, sz, nullptr, 1); launches++; }
        const int nsz = sz + 1;
        LQ_CUDA(cudaMemcpyAsync(tv.size, &nsz, sizeof(int), cudaMemcpyHostToDevice, st));
        LQ_CUDA(cudaStreamSynchronize(st));
        size_ub = nsz; size_ub_exact = nsz; last_stats.tree_size = nsz;
        return LQ_OK;
    }

    int mark_ignore(const int* ids, int count, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (count <= 0) return LQ_OK;
        DevBuf<int> d;
        std::vector<int> v(ids, ids + count);
        if (int r = upload(d, v)) return r;
        mark_ignore_kernel<T><<<(count + 127) / 128, 128, 0, st>>>(tv, d.p, count); launches++;
        LQ_CUDA(cudaGetLastError());
        LQ_CUDA(cudaStreamSynchronize(st));
        return LQ_OK;
    }

    int climb(int node, int* ids, int max_len, int* len, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        int sz = 0;
        LQ_CUDA(cudaStreamSynchronize(st));
        LQ_CUDA(cudaMemcpy(&sz, tv.size, sizeof(int), cudaMemcpyDeviceToHost));
        if (node < 0 || node >= sz) LQ_FAIL(LQ_ERR_ARG, "The given ID, %d, doesn't exist.", node);      // tree.py:109-110
        if (max_len < 1) LQ_FAIL(LQ_ERR_ARG, "max_len must be >= 1");
        DevBuf<int> d;
        LQ_CUDA(d.alloc((size_t)max_len + 1));
        climb_kernel<T><<<1, 1, 0, st>>>(tv, node, d.p, max_len, d.p + max_len); launches++;
        LQ_CUDA(cudaGetLastError());
        LQ_CUDA(cudaStreamSynchronize(st));
        int L = 0;
        LQ_CUDA(cudaMemcpy(&L, d.p + max_len, sizeof(int), cudaMemcpyDeviceToHost));
        if (L < 0) LQ_FAIL(LQ_ERR_STATE, "path from node %d is longer than %d", node, max_len);
        std::vector<int> tmp(L);
        LQ_CUDA(cudaMemcpy(tmp.data(), d.p, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost));
        for (int k = 0; k < L; k++) ids[k] = tmp[L - 1 - k];          // seed first, as tree.py:117
        *len = L;
        return LQ_OK;
    }

    // Plan read-out in one call (planner.py:266-281 / 313-324): [closest-to-guide ->] climb -> the edges of the path.
    // Two host round trips (ids, then edges) on persistent scratch instead of a dozen small synchronous copies and
    // three cudaMalloc/cudaFree pairs; what `update_plan` spends after its loop is part of the end-to-end figure.
    DevBuf<int> r_ids, r_len;
    DevBuf<T> r_x, r_u;
    int r_cap = 0;
    std::vector<T> r_hx, r_hu;
    int plan_readout(int node, const double* xg, const uint8_t* colmask, int max_len, int* ids, int* len,
                     double* hx, double* hu, int* elen, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (max_len < 1) LQ_FAIL(LQ_ERR_ARG, "max_len must be >= 1");
        if (r_cap < max_len) {
            LQ_CUDA(r_ids.alloc((size_t)max_len + 2)); LQ_CUDA(r_len.alloc(max_len));
            LQ_CUDA(r_x.alloc((size_t)max_len * H * N)); LQ_CUDA(r_u.alloc((size_t)max_len * H * M));
            r_cap = max_len;
        }
        int* d_node = r_ids.p + max_len + 1;
        int* d_plen = r_ids.p + max_len;
        if (node < 0) {                      // closest-to-guide first (planner.py:313-318)
            if (!xg) LQ_FAIL(LQ_ERR_ARG, "closest-to-guide needs the guide state");
            FullMetric<T, N> fm;
            double S[N * N];
            if (Lqr::CONST_METRIC) {
                for (int k = 0; k < N * N; k++) S[k] = Sconst[k];
            } else {
                double K[M * N], u0[M] = {0};
                if (int r = eval_lqr(1, xg, u0, S, K)) return r;
            }
            for (int i = 0; i < N; i++) {
                fm.x[i] = (T)xg[i];
                for (int j = 0; j < N; j++) fm.S[i * N + j] = (colmask && colmask[j]) ? T(0) : (T)S[i * N + j];
            }
            fm.angle_mask = md.angle_mask;
            closest_kernel<T, N><<<CLOSEST_CTAS, CLOSEST_BLOCK, 0, st>>>(t_state.p, cap, tv.size, fm, d_node, pcost_of(0), pidx_of(0), (unsigned*)(w_misc.p + 4)); launches++;
        } else {
            LQ_CUDA(cudaMemcpyAsync(d_node, &node, sizeof(int), cudaMemcpyHostToDevice, st));
        }
        climb_from_kernel<T><<<1, 1, 0, st>>>(tv, d_node, r_ids.p, max_len, d_plen); launches++;
        LQ_CUDA(cudaGetLastError());
        std::vector<int> tmp((size_t)max_len + 2);
        LQ_CUDA(cudaMemcpyAsync(tmp.data(), r_ids.p, ((size_t)max_len + 2) * sizeof(int), cudaMemcpyDeviceToHost, st));
        LQ_CUDA(cudaStreamSynchronize(st));
        const int L = tmp[max_len];
        if (L < 0) LQ_FAIL(LQ_ERR_STATE, "path from node %d is longer than %d", tmp[max_len + 1], max_len);
        for (int k = 0; k < L; k++) ids[k] = tmp[L - 1 - k];          // seed first, as tree.py:117
        *len = L;
        if (L == 0 || !hx) return LQ_OK;
        // edges of the path, leaf-first on the device, reordered seed-first on the way out
        gather_edges_kernel<T><<<L, 128, 0, st>>>(tv, r_ids.p, L, last_my_seg, H, H * N, H * M, r_x.p, r_u.p, r_len.p); launches++;
        LQ_CUDA(cudaGetLastError());
        r_hx.resize((size_t)L * H * N); r_hu.resize((size_t)L * H * M);
        std::vector<int> el(L);
        LQ_CUDA(cudaMemcpyAsync(r_hx.data(), r_x.p, r_hx.size() * sizeof(T), cudaMemcpyDeviceToHost, st));
        LQ_CUDA(cudaMemcpyAsync(r_hu.data(), r_u.p, r_hu.size() * sizeof(T), cudaMemcpyDeviceToHost, st));
        LQ_CUDA(cudaMemcpyAsync(el.data(), r_len.p, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost, st));
        LQ_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < L; k++) {
            const int src = L - 1 - k;
            for (int q = 0; q < H * N; q++) hx[(size_t)k * H * N + q] = (double)r_hx[(size_t)src * H * N + q];
            for (int q = 0; q < H * M; q++) hu[(size_t)k * H * M + q] = (double)r_hu[(size_t)src * H * M + q];
            elen[k] = el[src];
        }
        return LQ_OK;
    }

    // ---- read-back --------------------------------------------------------------
    int download_nodes(int first, int count, double* state, int* parent, int* cost, int* elen, uint8_t* ign,
                       uint8_t* owner, double* K, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        LQ_CUDA(cudaStreamSynchronize(st));
        int sz = 0;
        LQ_CUDA(cudaMemcpy(&sz, tv.size, sizeof(int), cudaMemcpyDeviceToHost));
        if (first < 0 || count < 0 || first + count > sz) LQ_FAIL(LQ_ERR_ARG, "node range [%d,%d) outside tree of %d", first, first + count, sz);
        if (count == 0) return LQ_OK;
        if (state) {
            std::vector<T> b((size_t)N * count);
            for (int d = 0; d < N; d++)
                LQ_CUDA(cudaMemcpy(b.data() + (size_t)d * count, t_state.p + (size_t)d * cap + first, count * sizeof(T), cudaMemcpyDeviceToHost));
            for (int i = 0; i < count; i++)
                for (int d = 0; d < N; d++) state[(size_t)i * N + d] = (double)b[(size_t)d * count + i];
        }
        if (K) {
            std::vector<T> b((size_t)M * N * count);
            for (int k = 0; k < M * N; k++)
                LQ_CUDA(cudaMemcpy(b.data() + (size_t)k * count, t_K.p + (size_t)k * cap + first, count * sizeof(T), cudaMemcpyDeviceToHost));
            for (int i = 0; i < count; i++)
                for (int k = 0; k < M * N; k++) K[(size_t)i * M * N + k] = (double)b[(size_t)k * count + i];
        }
        if (parent) LQ_CUDA(cudaMemcpy(parent, t_parent.p + first, count * sizeof(int), cudaMemcpyDeviceToHost));
        if (cost) LQ_CUDA(cudaMemcpy(cost, t_cost.p + first, count * sizeof(int), cudaMemcpyDeviceToHost));
        if (elen) LQ_CUDA(cudaMemcpy(elen, t_elen.p + first, count * sizeof(int), cudaMemcpyDeviceToHost));
        if (ign) LQ_CUDA(cudaMemcpy(ign, tv.ignore[ign_copy] + first, count, cudaMemcpyDeviceToHost));
        if (owner) LQ_CUDA(cudaMemcpy(owner, t_owner.p + first, count, cudaMemcpyDeviceToHost));
        return LQ_OK;
    }

    int download_edges(const int* ids, int count, double* hx, double* hu, int* hlen, cudaStream_t st) override {
        LQ_CUDA(cudaSetDevice(device));
        if (count <= 0) return LQ_OK;
        int sz = 0;
        LQ_CUDA(cudaStreamSynchronize(st));
        LQ_CUDA(cudaMemcpy(&sz, tv.size, sizeof(int), cudaMemcpyDeviceToHost));
        for (int i = 0; i < count; i++)
            if (ids[i] < 0 || ids[i] >= sz) LQ_FAIL(LQ_ERR_ARG, "The given ID, %d, doesn't exist.", ids[i]);
        DevBuf<int> d;
        std::vector<int> v(ids, ids + count);
        if (int r = upload(d, v)) return r;
        DevBuf<T> ox, ou;
        DevBuf<int> ol;
        LQ_CUDA(ox.alloc((size_t)count * H * N)); LQ_CUDA(ou.alloc((size_t)count * H * M)); LQ_CUDA(ol.alloc(count));
        gather_edges_kernel<T><<<count, 128, 0, st>>>(tv, d.p, count, last_my_seg, H, H * N, H * M, ox.p, ou.p, ol.p); launches++;
        LQ_CUDA(cudaGetLastError());
        std::vector<T> bx((size_t)count * H * N), bu((size_t)count * H * M);
        LQ_CUDA(cudaStreamSynchronize(st));
        LQ_CUDA(cudaMemcpy(bx.data(), ox.p, bx.size() * sizeof(T), cudaMemcpyDeviceToHost));
        LQ_CUDA(cudaMemcpy(bu.data(), ou.p, bu.size() * sizeof(T), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < bx.size(); k++) hx[k] = (double)bx[k];
        for (size_t k = 0; k < bu.size(); k++) hu[k] = (double)bu[k];
        if (hlen) LQ_CUDA(cudaMemcpy(hlen, ol.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost));
        return LQ_OK;
    }

    // ---- the callbacks, batched ---------------------------------------------------
    int eval_generic(int what, int count, const double* a, int na, const double* b, int nb, double dt,
                     double* o1, int n1, double* o2, int n2, uint8_t* ob) {
        LQ_CUDA(cudaSetDevice(device));
        if (count <= 0) return LQ_OK;
        std::vector<T> ha((size_t)count * na), hb((size_t)count * nb);
        for (size_t k = 0; k < ha.size(); k++) ha[k] = (T)a[k];
        for (size_t k = 0; k < hb.size(); k++) hb[k] = (T)b[k];
        DevBuf<T> da, db, d1, d2;
        DevBuf<uint8_t> dob;
        if (int r = upload(da, ha)) return r;
        if (int r = upload(db, hb)) return r;
        LQ_CUDA(d1.alloc((size_t)count * std::max(1, n1))); LQ_CUDA(d2.alloc((size_t)count * std::max(1, n2)));
        LQ_CUDA(dob.alloc(count));
        eval_kernel<T, Sys, Lqr><<<(count + 63) / 64, 64>>>(md, what, count, da.p, db.p, (T)dt, d1.p, d2.p, dob.p);
        launches++;
        LQ_CUDA(cudaGetLastError());
        LQ_CUDA(cudaDeviceSynchronize());
        if (o1 && n1) {
            std::vector<T> h((size_t)count * n1);
            LQ_CUDA(cudaMemcpy(h.data(), d1.p, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
            for (size_t k = 0; k < h.size(); k++) o1[k] = (double)h[k];
        }
        if (o2 && n2) {
            std::vector<T> h((size_t)count * n2);
            LQ_CUDA(cudaMemcpy(h.data(), d2.p, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
            for (size_t k = 0; k < h.size(); k++) o2[k] = (double)h[k];
        }
        if (ob) LQ_CUDA(cudaMemcpy(ob, dob.p, count, cudaMemcpyDeviceToHost));
        return LQ_OK;
    }
    int eval_dynamics(int count, const double* x, const double* u, double dt, double* xn) override {
        return eval_generic(0, count, x, N, u, M, dt, xn, N, nullptr, 0, nullptr);
    }
    int eval_lqr(int count, const double* x, const double* u, double* S, double* K) override {
        return eval_generic(1, count, x, N, u, M, 0.0, S, N * N, K, M * N, nullptr);
    }
    int eval_erf(int count, const double* xg, const double* x, double* e) override {
        return eval_generic(2, count, xg, N, x, N, 0.0, e, N, nullptr, 0, nullptr);
    }
    int eval_feasible(int count, const double* x, const double* u, uint8_t* ok) override {
        return eval_generic(3, count, x, N, u, M, 0.0, nullptr, 0, nullptr, 0, ok);
    }

    int set_tuning(int lanes, int nq) override {
        if (lanes == 1 || lanes == 8 || lanes == 32) steer_lanes = lanes;
        else if (lanes != 0) LQ_FAIL(LQ_ERR_ARG, "steer lanes must be 0 (keep), 1, 8 or 32");
        near_q = nq;
        return LQ_OK;
    }

    int prefetch_setup() {
        if (aux) return LQ_OK;
        LQ_CUDA(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
        for (auto& e : ev_sampled) LQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ev_consumed) LQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return LQ_OK;
    }
    int set_option(int option, int value) override {
        LQ_CUDA(cudaSetDevice(device));
        switch (option) {
        case LQ_OPT_NEAREST_KERNEL:
            if (value < LQ_NEAREST_SCALAR || value > LQ_NEAREST_PREFIX) LQ_FAIL(LQ_ERR_ARG, "unknown nearest kernel %d", value);
            near_kernel = value; return LQ_OK;
        case LQ_OPT_NEAREST_CTAS_PER_SM:
            if (value < 0 || value > 64) LQ_FAIL(LQ_ERR_ARG, "CTAs per SM must be in 0 (auto)..64");
            near_ctas_per_sm = value; return LQ_OK;
        case LQ_OPT_SAMPLE_PREFETCH:
            prefetch_on = value != 0; return LQ_OK;
        case LQ_OPT_SPECULATIVE_SCAN: spec_scan = value != 0; return LQ_OK;
        case LQ_OPT_FORCE_RESCAN: force_rescan = value != 0; return LQ_OK;
        case LQ_OPT_STAGE_TABLES: stage_tables_on = value != 0; return LQ_OK;
        case LQ_OPT_PDL: pdl_on = value != 0; return LQ_OK;
        case LQ_OPT_SAMPLE_AHEAD: sample_ahead = value != 0; return LQ_OK;
        case LQ_OPT_WHITEN_METRIC:
            if (q_head != q_tail) LQ_FAIL(LQ_ERR_STATE, "waves are in flight");
            whiten_opt = value != 0; return LQ_OK;
        case LQ_OPT_NEAREST_GRID:
            if (value < 0 || value > 1) LQ_FAIL(LQ_ERR_ARG, "indexed search must be 0 or 1");
            if (value != grid_opt) {
                if (q_head != q_tail) LQ_FAIL(LQ_ERR_STATE, "waves are in flight");
                grid_opt = value;
                tv.grid.on = grid_active() ? 1 : 0;
                if (tv.grid.on) grid_dirty = true;       // nothing was linked while it was off
            }
            return LQ_OK;
        case LQ_OPT_GRID_MIN_KPAIRS:
            if (value < 0) LQ_FAIL(LQ_ERR_ARG, "threshold must be >= 0");
            grid_min_kpairs = value; return LQ_OK;
        case LQ_OPT_STEER_BLOCK:
            if (value != 0 && value != 32 && value != 128) LQ_FAIL(LQ_ERR_ARG, "rollout CTA size must be 0 (auto), 32 or 128");
            steer_block = value; return LQ_OK;
        default: LQ_FAIL(LQ_ERR_ARG, "unknown option %d", option);
        }
    }
    int get_option(int option, int* value) override {
        switch (option) {
        case LQ_OPT_NEAREST_KERNEL: *value = near_kernel; return LQ_OK;
        case LQ_OPT_NEAREST_CTAS_PER_SM: *value = near_ctas_per_sm; return LQ_OK;
        case LQ_OPT_SAMPLE_PREFETCH: *value = prefetch_on ? 1 : 0; return LQ_OK;
        case LQ_OPT_SPECULATIVE_SCAN: *value = spec_scan; return LQ_OK;
        case LQ_OPT_FORCE_RESCAN: *value = force_rescan; return LQ_OK;
        case LQ_OPT_STAGE_TABLES: *value = stage_tables_on; return LQ_OK;
        case LQ_OPT_STEER_BLOCK: *value = steer_block; return LQ_OK;
        case LQ_OPT_PDL: *value = pdl_on; return LQ_OK;
        case LQ_OPT_SAMPLE_AHEAD: *value = sample_ahead; return LQ_OK;
        case LQ_OPT_WHITEN_METRIC: *value = whiten_active() ? 1 : 0; return LQ_OK;
        case LQ_OPT_NEAREST_GRID: *value = grid_active() ? 1 : 0; return LQ_OK;
        case LQ_OPT_GRID_MIN_KPAIRS: *value = grid_min_kpairs; return LQ_OK;
        default: LQ_FAIL(LQ_ERR_ARG, "unknown option %d", option);
        }
    }

    int set_profiling(int on) override {
        LQ_CUDA(cudaSetDevice(device));
        if (on && !pev[0][0]) for (auto& r : pev) for (auto& e : r) LQ_CUDA(cudaEventCreate(&e));
        prof_on = on != 0;
        for (int k = 0; k < RING; k++) prof_open[k] = false;
        return LQ_OK;
    }
    int profile_read(double* ms, double* work, int reset) override {
        if (ms) for (int k = 0; k < 5; k++) ms[k] = prof_ms[k];
        if (work) for (int k = 0; k < 4; k++) work[k] = prof_work[k];
        if (reset) { for (auto& v : prof_ms) v = 0; for (auto& v : prof_work) v = 0; }
        return LQ_OK;
    }
};

template <typename T, class Sys, class Lqr>
lq_engine* make_engine(const lq_config* cfg, std::string& err, int* rc) {
    auto* e = new EngineImpl<T, Sys, Lqr>();
    *rc = e->init(cfg);
    if (*rc != LQ_OK) { err = e->err; delete e; return nullptr; }
    return e;
}

}  // namespace lq
