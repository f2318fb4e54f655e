// K2: Planner._steer(ID, xtar, force_arrive=False) for a whole wave
// (reference lqrrt/planner.py:354-438; exact control flow SURVEY.md section 3.2),
// fused with the lqr provider, the dynamics step and is_feasible -- and, inside a wave, with the
// selection of the nearest node from K1's partial minima (planner.py:239-247) in its prologue and the
// exchange of the result with every replica of the tree in its epilogue.
//
// One sample per group of LANES threads.  All lanes of a group carry the same
// state/gain in registers and execute the same instruction stream (results are
// bitwise identical across lanes); only the obstacle tables of is_feasible are
// split across the lanes, and the verdict is combined with a group ballot.
// LANES = 1 is thread-per-sample; LANES = 32 is warp-per-sample for heavy
// collision fields.  Rollout is latency/issue bound: tiny matrices, no tensor
// cores; small obstacle tables are staged in shared memory once per CTA (models.cuh stage_tables).
#pragma once
#include <cooperative_groups.h>
#include "models.cuh"
#include "nearest.cuh"
#include "wave.cuh"

namespace lq {
namespace cg = cooperative_groups;

template <typename T, class Sys, class Lqr> struct Model {
    Sys sys;
    Lqr lqr;
    FeasSet<T> feas;
    uint32_t angle_mask;
    int H;                       // horizon_iters (planner.py:549)
    T dt;
    T tol[LQ_MAX_N];             // error_tol (planner.py:532-536)
    T goal_lo[LQ_MAX_N], goal_hi[LQ_MAX_N];   // goal region, open box (planner.py:442-447, 482-486)
    double FPR;                  // int(FPR * len) is evaluated in float64 as Python does (planner.py:394-395)
    // adaptive-horizon heuristic (planner.py:418-425, 538-547); hfactor == 0: fixed horizon H
    int hfactor, hlo, hhi;
    const int* horizon;          // device scalar: planner-level horizon_iters as of wave start
};

// int(np.clip(h / f, lo, hi)) and int(np.clip(f * h, lo, hi)) of planner.py:421,424 (float64, truncation)
__host__ __device__ inline int horizon_shrink(int h, int f, int lo, int hi) {
    double v = (double)h / (double)f;
    v = v < lo ? lo : (v > hi ? hi : v);
    return (int)v;
}
__host__ __device__ inline int horizon_grow(int h, int f, int lo, int hi) {
    long long v = (long long)f * h;
    v = v < lo ? lo : (v > hi ? hi : v);
    return (int)v;
}
// a rollout's effect on the planner-level horizon: low 7 bits = number of doublings, bit 7 = halved
__host__ __device__ inline int horizon_apply(int h, unsigned vote, int f, int lo, int hi) {
    for (unsigned k = 0; k < (vote & 0x7fu); k++) h = horizon_grow(h, f, lo, hi);
    if (vote & 0x80u) h = horizon_shrink(h, f, lo, hi);
    return h;
}

template <typename T, class Sys, class Lqr>
__device__ __forceinline__ int fpr_keep(const Model<T, Sys, Lqr>& md, int len) { return (int)(md.FPR * (double)len); }

// Shared-memory ordering inside the group of LANES threads that share one sample.
template <int LANES>
struct GroupSync {
    __device__ __forceinline__ void operator()() const {
        if (LANES > 1) cg::tiled_partition<LANES>(cg::this_thread_block()).sync();
    }
    __device__ __forceinline__ bool all(bool p) const {
        if (LANES > 1) return cg::tiled_partition<LANES>(cg::this_thread_block()).all(p);
        return p;
    }
};

// K = lqr(x, u).K, cooperatively when the provider supports it (FD-Jacobian + Riccati) and the group has lanes.
// have_prev: the group's rollout-long note that its shared-memory slice holds the previous solve (FD-Riccati: a gain is
// kept while the Jacobian does not change, models.cuh solve_coop)
template <typename T, class Sys, class Lqr, int LANES>
__device__ __forceinline__ void group_gain(const Model<T, Sys, Lqr>& md, const T* x, const T* u, T* K, T s, T c,
                                           T* coop, int lane, int* have_prev = nullptr) {
    if constexpr (Lqr::COOP && LANES > 1) {
        md.lqr.gain_coop(md.sys, md.dt, x, u, K, coop, lane, LANES, GroupSync<LANES>(), have_prev);
    } else {
        md.lqr.gain_sc(md.sys, md.dt, x, u, K, s, c);
    }
}

template <int LANES>
__device__ __forceinline__ bool group_all(bool ok) {
    if (LANES == 1) return ok;
    // control flow is uniform inside a group, so the tile collective is convergent
    return cg::tiled_partition<LANES>(cg::this_thread_block()).all(ok);
}

// ---- the rollout itself (planner.py:373-438) --------------------------------------------------------------
// In: x = the node's state, xt = the target, K = the node's stored gain.  The recorded edge goes to
// xrow [H][N] / urow [H][M] (lane 0 writes).  Out: len = kept length after the FPR cut, vote = the
// adaptive-horizon vote, and -- when len > 0 -- x / K hold the new node: the last KEPT state and
// lqr(xnew, u_seq[-1]).K (planner.py:256-257).
template <typename T, class Sys, class Lqr, int LANES>
__device__ __forceinline__ void rollout(const Model<T, Sys, Lqr>& md, const FeasTab<T>* tabs,
                                        T (&x)[Sys::N], const T (&xt)[Sys::N], T (&K)[Sys::M * Sys::N],
                                        T* __restrict__ xrow, T* __restrict__ urow, int lane, T* coop,
                                        int& len_out, unsigned& vote_out) {
    constexpr int N = Sys::N, M = Sys::M;
    // sin/cos of the heading-like dim are evaluated once per state and shared by erf, dynamics,
    // the hull rotation in is_feasible and the gain (they were ~half of a step's instructions)
    constexpr int HD = Sys::HEADING;
    T sh = T(0), ch = T(1), sg = T(0), cg = T(1);
    if (HD >= 0) { sincos_t(x[HD < 0 ? 0 : HD], &sh, &ch); sincos_t(xt[HD < 0 ? 0 : HD], &sg, &cg); }

    int len = 0, nrec = 0, i = 0;
    int have_prev = 0;           // nothing of this rollout in the provider's shared-memory slice yet
    // every rollout of a wave starts from the same planner-level horizon (frozen at wave start)
    const bool adaptive = md.hfactor != 0;
    int hor = adaptive ? *md.horizon : md.H;
    unsigned vote = 0;
    T last_emag[N];
#pragma unroll
    for (int d = 0; d < N; d++) last_emag[d] = Num<T>::inf();

    while (true) {
        T e[N], u[M], uw[M], xn[N];
        erf_eval_sc<T, N, HD>(md.angle_mask, xt, x, e, sg, cg, sh, ch);   // planner.py:386
#pragma unroll
        for (int r = 0; r < M; r++) {                            // u = K e, planner.py:387
            T a = T(0);
#pragma unroll
            for (int d = 0; d < N; d++) a += K[r * N + d] * e[d];
            u[r] = a; uw[r] = a;
        }
        md.sys.step_sc(x, uw, md.dt, xn, sh, ch);                // planner.py:390 (uw = the copy dynamics may clip)
        T sn = T(0), cn = T(1);
        if (HD >= 0) sincos_t(xn[HD < 0 ? 0 : HD], &sn, &cn);
        const bool ok = group_all<LANES>(feasible_eval<T, N, M>(md.feas, xn, u, lane, LANES, HD, sn, cn, tabs));   // planner.py:393
        if (!ok) { len = fpr_keep(md, nrec); break; }            // planner.py:394-396
        i += 1;                                                  // planner.py:414
        bool conv = true;
#pragma unroll
        for (int d = 0; d < N; d++) conv = conv && (abs_t(e[d]) <= md.tol[d]);
        if (adaptive) {                                          // planner.py:418-425
            bool stalled = true;
#pragma unroll
            for (int d = 0; d < N; d++) stalled = stalled && (abs_t(e[d]) >= last_emag[d]);
            if (stalled) { len = 0; vote |= 0x80u; break; }      // no error shrank: drop the edge, halve the horizon
            if (i == hor) { hor = horizon_grow(hor, md.hfactor, md.hlo, md.hhi); vote += 1; }
#pragma unroll
            for (int d = 0; d < N; d++) last_emag[d] = abs_t(e[d]);
        }
        if (i > hor || conv) { len = nrec; break; }              // planner.py:428-429
#pragma unroll
        for (int d = 0; d < N; d++) x[d] = xn[d];
        sh = sn; ch = cn;
        if (lane == 0) {                                         // planner.py:432-433
#pragma unroll
            for (int d = 0; d < N; d++) xrow[nrec * N + d] = x[d];
#pragma unroll
            for (int r = 0; r < M; r++) urow[nrec * M + r] = u[r];
        }
        nrec += 1;
        group_gain<T, Sys, Lqr, LANES>(md, x, u, K, sh, ch, coop, lane, &have_prev);   // planner.py:436
    }

    // The node is the last KEPT state and its gain is lqr(xnew, u_seq[-1]).K (planner.py:256-257).  Without an
    // FPR cut, x / K already hold exactly that; after a cut (0 < len < nrec) the kept tail is re-read from
    // the group's own edge rows (written by lane 0) and the gain recomputed by the whole group.
    if (len > 0 && len < nrec) {
        GroupSync<LANES>()();
        T uk[M];
#pragma unroll
        for (int d = 0; d < N; d++) x[d] = xrow[(len - 1) * N + d];
#pragma unroll
        for (int r = 0; r < M; r++) uk[r] = urow[(len - 1) * M + r];
        T s2 = T(0), c2 = T(1);
        if (HD >= 0) sincos_t(x[HD < 0 ? 0 : HD], &s2, &c2);
        group_gain<T, Sys, Lqr, LANES>(md, x, uk, K, s2, c2, coop, lane, &have_prev);
    }
    len_out = len;
    vote_out = vote;
}

// dynamic shared memory of the rollout kernels: [cooperative-provider slices][staged obstacle tables]
template <typename T, class Lqr, int LANES, int BLOCK>
__host__ __device__ constexpr size_t steer_coop_bytes() {
    return (Lqr::COOP && LANES > 1) ? align16((size_t)(BLOCK / LANES) * Lqr::COOP_SCALARS * sizeof(T)) : 0;
}

// ---- stand-alone rollouts: lq_steer (W given (node, target) pairs) ---------------------------------------
// Outputs per pair j: xseq [j][H][N], useq [j][H][M] recorded edge (sample-major), len [j].
template <typename T, class Sys, class Lqr, int LANES, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
steer_kernel(Model<T, Sys, Lqr> md, StagePlan stg,
             const T* __restrict__ tstate, const T* __restrict__ tK, int cap,
             const int* __restrict__ ids, const T* __restrict__ targets, int tstride, int W,
             T* __restrict__ xseq, T* __restrict__ useq, int* __restrict__ len_out) {
    constexpr int N = Sys::N, M = Sys::M;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ FeasTab<T> tabs[LQ_MAX_PRIMS];
    __shared__ __align__(8) unsigned long long stage_bar;
    T* coop = reinterpret_cast<T*>(dyn_smem) + (size_t)(threadIdx.x / LANES) * Lqr::COOP_SCALARS;
    stage_tables(md.feas, stg, tabs, dyn_smem + steer_coop_bytes<T, Lqr, LANES, BLOCK>(), &stage_bar);
    const int j = (blockIdx.x * BLOCK + threadIdx.x) / LANES;
    const int lane = threadIdx.x % LANES;
    if (j >= W) return;                    // whole groups leave together (BLOCK % LANES == 0)
    const int id = ids[j];
    T x[N], xt[N], K[M * N];
#pragma unroll
    for (int d = 0; d < N; d++) { x[d] = tstate[(size_t)d * cap + id]; xt[d] = targets[(size_t)d * tstride + j]; }
#pragma unroll
    for (int k = 0; k < M * N; k++) K[k] = tK[(size_t)k * cap + id];
    int len; unsigned vote;
    rollout<T, Sys, Lqr, LANES>(md, tabs, x, xt, K, xseq + (size_t)j * md.H * N, useq + (size_t)j * md.H * M, lane, coop, len, vote);
    if (lane == 0) len_out[j] = len;
}

// ---- nearest selection in the rollout kernel's prologue ----------------------------------------------------
template <typename T>
struct PartialsIn { const T* pcost; const int* pidx; int pstride; int nslots; };

// How K1 measured distance for this engine: the rare per-query rescan below must reproduce it bit for bit.
template <typename T, int N>
struct NearMetric {
    DiagMetric<T, N> dmet;
    const T* Spacked;        // per-query packed S (dense metrics), [NS][sstride]
    int diag;                // 1: diagonal metric on the states; 2: identity metric on the whitened copy (dmet = identity)
    const T* zstate;         // diag == 2: z = U v per node, [N][cap]
    const T* U;
};

// One query's brute-force scan by its own group of LANES threads: the exact arithmetic of the K1 kernels
// (nearest.cuh), lexicographic (cost, index) minimum.  Used when the merged winner of the speculative scans
// turned out to be ignored (a commit that ran concurrently with the old-tree scan marked it) or when no
// un-ignored node is left (planner.py:244-245: plain argmin) -- rare by construction.
template <typename T, int N, int LANES>
__device__ __noinline__ int rescan_nearest(const T* __restrict__ tstate, int cap, const uint8_t* __restrict__ ignore,
                                           int size, bool mask_on, const NearMetric<T, N>& nm, const T (&xq)[N],
                                           int j, int sstride, int lane) {
    constexpr bool EXACT = sizeof(T) == 8;
    constexpr int NS = N * (N + 1) / 2;
    const uint32_t mask = nm.dmet.angle_mask;
    T q[N], S[NS];
    if (nm.diag == 2) tstate = nm.zstate;            // the scan searched the whitened copy: so does this
    if (nm.diag) {
        T xw[N];
        if (nm.diag == 2) {
            whiten_apply<T, N>(nm.U, xq, xw);
        } else {
#pragma unroll
            for (int d = 0; d < N; d++) xw[d] = xq[d];
        }
#pragma unroll
        for (int d = 0; d < N; d++) {
            T v = xw[d];
            if (!EXACT) { if (mask & (1u << d)) v = reduce_period(v); v = prescale(v, nm.dmet.w[d]); }
            q[d] = v;
        }
    } else {
#pragma unroll
        for (int d = 0; d < N; d++) q[d] = xq[d];
#pragma unroll
        for (int s = 0; s < NS; s++) S[s] = nm.Spacked[(size_t)s * sstride + j];
    }
    T best = Num<T>::inf();
    int b = -1;
    for (int node = lane; node < size; node += LANES) {
        if (mask_on && ignore[node]) continue;
        T c;
        if (nm.diag) {
            c = T(0);
#pragma unroll
            for (int d = 0; d < N; d++) {
                T a = tstate[(size_t)d * cap + node];
                if (EXACT) {
                    T e = q[d] - a;
                    if (mask & (1u << d)) e = reduce_period(e);
                    c = diag_term_exact(c, e, nm.dmet.sdiag[d]);
                } else {
                    if (mask & (1u << d)) a = reduce_period(a);
                    a = prescale(a, nm.dmet.w[d]);
                    T e = q[d] - a;
                    if (mask & (1u << d)) e = diag_wrap_scaled(e, nm.dmet.period[d]);
                    c = fma(e, e, c);
                }
            }
        } else {
            T e[N];
#pragma unroll
            for (int d = 0; d < N; d++) {
                e[d] = q[d] - tstate[(size_t)d * cap + node];
                if (mask & (1u << d)) e[d] = reduce_period(e[d]);
            }
            c = dense_cost<T, N>(e, S);
        }
        if (c < best) { best = c; b = node; }            // ascending per lane: the first minimum of this lane
    }
    if (LANES > 1) {
        auto tile = cg::tiled_partition<LANES>(cg::this_thread_block());
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            const T c2 = tile.shfl_xor(best, o);
            const int i2 = tile.shfl_xor(b, o);
            merge_candidate(best, b, c2, i2);
        }
    }
    return b;
}

// ---- one wave's rollouts (the wave pipeline's K2) ----------------------------------------------------------
// For every sample j of this rank's slice (count read from the wave's control slot):
//   prologue  merge K1's partial minima (old-tree scan + delta scan slots) into the nearest node, check
//             that it is still un-ignored (rescan otherwise);
//   rollout   as above, edge into the per-sample scratch rows xseq / useq;
//   epilogue  meta word + (if a node came out) parent and record row into segment `rank` of every
//             destination in `dst`; in peer mode the last CTA raises this rank's arrival flag on every peer.
template <typename T, class Sys, class Lqr, int LANES, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
steer_x_kernel(Model<T, Sys, Lqr> md, StagePlan stg,
               const T* __restrict__ tstate, const T* __restrict__ tK, int cap,
               const uint8_t* __restrict__ ignore, const int* __restrict__ d_size, const int* __restrict__ d_nignored,
               WaveCtl* __restrict__ ctl, int slot, PartialsIn<T> pin, NearMetric<T, Sys::N> nm,
               const T* __restrict__ samples, int sstride, const int* __restrict__ tries_used,
               int use_ignore, int force_rescan,
               T* __restrict__ xseq, T* __restrict__ useq,
               ExDst dst, int Wc, int host_flags, const int* __restrict__ h_stop, int done_slot) {
    constexpr int N = Sys::N, M = Sys::M;
    constexpr int PER16 = 16 / (int)sizeof(T);
    constexpr int RW = (N + M * N + PER16 - 1) / PER16 * PER16;
    constexpr int NCH = RW / PER16;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ FeasTab<T> tabs[LQ_MAX_PRIMS];
    __shared__ __align__(8) unsigned long long stage_bar;
    __shared__ bool s_last;
    T* coop = reinterpret_cast<T*>(dyn_smem) + (size_t)(threadIdx.x / LANES) * Lqr::COOP_SCALARS;

    const int count = ctl->ring[slot].count;
    const int j = (blockIdx.x * BLOCK + threadIdx.x) / LANES;
    const int lane = threadIdx.x % LANES;
    const bool active = j < count;                      // whole groups (BLOCK % LANES == 0)

    if (blockIdx.x == 0 && threadIdx.x == 0) {          // this rank's segment header, on every replica
        int flags = host_flags;
        if (h_stop) flags |= *(const volatile int*)h_stop;      // the host's stop request (mapped host memory)
        for (int p = 0; p < dst.n; p++) {
            int* h = (int*)dst.seg[p];
            h[HDR_SAMPLED] = count; h[HDR_FLAGS] = flags;
        }
    }
    // nothing to roll out in this CTA: skip the table staging
    const bool cta_active = blockIdx.x * (BLOCK / LANES) < count;
    if (cta_active) stage_tables(md.feas, stg, tabs, dyn_smem + steer_coop_bytes<T, Lqr, LANES, BLOCK>(), &stage_bar);

    pdl_trigger();                      // the commit kernel may be launched; it waits for this grid to finish
    pdl_wait();                         // the scan in front of this kernel has finished: its partial minima are visible
    if (active) {
        T x[N], xt[N], K[M * N];
#pragma unroll
        for (int d = 0; d < N; d++) xt[d] = samples[(size_t)d * sstride + j];
        // ---- nearest node: lexicographic merge of the partial minima, lanes split the slots
        int id;
        {
            T best = Num<T>::inf();
            int b = -1;
            for (int sp = lane; sp < pin.nslots; sp += LANES)
                merge_candidate(best, b, pin.pcost[(size_t)sp * pin.pstride + j], pin.pidx[(size_t)sp * pin.pstride + j]);
            if (LANES > 1) {
                auto tile = cg::tiled_partition<LANES>(cg::this_thread_block());
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) {
                    const T c2 = tile.shfl_xor(best, o);
                    const int i2 = tile.shfl_xor(b, o);
                    merge_candidate(best, b, c2, i2);
                }
            }
            // The scans treat ignored nodes as absent.  A winner that is ignored NOW was marked by a commit
            // that overlapped the old-tree scan; no winner means every node is ignored.  Either way this
            // query is searched again against the tree as it is (marks only ever get added, so a winner
            // that is still un-ignored is the winner of the up-to-date candidate set too).
            if (force_rescan || b < 0 || (use_ignore && ignore[b])) {
                const int size = *d_size;
                const bool mask_on = use_ignore && (*d_nignored < size);
                b = rescan_nearest<T, N, LANES>(tstate, cap, ignore, size, mask_on, nm, xt, j, sstride, lane);
            }
            id = b < 0 ? 0 : b;
        }
#pragma unroll
        for (int d = 0; d < N; d++) x[d] = tstate[(size_t)d * cap + id];
#pragma unroll
        for (int k = 0; k < M * N; k++) K[k] = tK[(size_t)k * cap + id];

        int len; unsigned vote;
        rollout<T, Sys, Lqr, LANES>(md, tabs, x, xt, K, xseq + (size_t)j * md.H * N, useq + (size_t)j * md.H * M, lane, coop, len, vote);

        // ---- epilogue: results to every replica
        const unsigned meta = meta_pack(len, vote, tries_used ? tries_used[j] : 0);
        for (int p = 0; p < dst.n; p++) {
            const SegView<T> sv = seg_view<T>(dst.seg[p], Wc, N, M);
            if (lane == 0) { sv.meta[j] = meta; if (len > 0) sv.parent[j] = id; }
            if (len > 0) {
                T* row = sv.rec + (size_t)j * RW;
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    if (c % LANES == lane) {
                        T v[PER16];
#pragma unroll
                        for (int t = 0; t < PER16; t++) {
                            const int e = c * PER16 + t;          // compile-time after unrolling: registers only
                            v[t] = e < N ? x[e < N ? e : 0] : (e < N + M * N ? K[(e >= N && e < N + M * N) ? e - N : 0] : T(0));
                        }
                        if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(row + c * PER16) = make_float4(v[0], v[1], v[2], v[3]);
                        else *reinterpret_cast<double2*>(row + c * PER16) = make_double2(v[0], v[1]);
                    }
                }
            }
        }
    }

    if (dst.flag[0]) {
        // fused exchange: my stores are visible system-wide before I count myself done; the last CTA of the
        // grid then raises this rank's arrival flag on every peer (after the data)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();                      // cumulative: covers the CTA's stores ordered before the barrier
            s_last = atomicAdd(&ctl->done_steer[done_slot], 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (s_last && threadIdx.x == 0) {
            __threadfence_system();
            for (int p = 0; p < dst.n; p++) *(volatile unsigned*)dst.flag[p] = dst.seq;
            ctl->done_steer[done_slot] = 0;
        }
    }
}

// Per-query metric S = lqr(x, 0).S for cooperative providers: LANES threads per query share one sweep
// (the thread-per-query metric_kernel in nearest.cuh serves the others).  Packed like metric_kernel.
template <typename T, class Sys, class Lqr, int LANES>
__global__ void __launch_bounds__(32)
metric_coop_kernel(Model<T, Sys, Lqr> md, const T* __restrict__ queries, int qstride, int W,
                   const int* __restrict__ d_count, T* __restrict__ Spacked) {
    constexpr int N = Sys::N, M = Sys::M;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    T* coop = reinterpret_cast<T*>(coop_raw) + (size_t)(threadIdx.x / LANES) * Lqr::SLOT_SCALARS;     // one solve per query: one slot
    const int j = (blockIdx.x * 32 + threadIdx.x) / LANES;
    const int lane = threadIdx.x % LANES;
    if (j >= (d_count ? min(*d_count, W) : W)) return;
    T x[N], u0[M];
#pragma unroll
    for (int d = 0; d < N; d++) x[d] = queries[(size_t)d * qstride + j];
#pragma unroll
    for (int r = 0; r < M; r++) u0[r] = T(0);                    // planner.py:348: lqr(x, zeros)
    if constexpr (Lqr::COOP) {
        md.lqr.solve_coop(md.sys, md.dt, x, u0, coop, lane, LANES, GroupSync<LANES>());
        const T* S = coop + Lqr::P_OFF;
        constexpr int NS = N * (N + 1) / 2;
        for (int e = lane; e < NS; e += LANES) {                 // packed index e -> (a, b)
            int a = 0, b = 0;
            if (e < N) { a = b = e; }
            else {
                int s = N;
                for (int aa = 0; aa < N; aa++)
                    for (int bb = aa + 1; bb < N; bb++, s++)
                        if (s == e) { a = aa; b = bb; }
            }
            Spacked[(size_t)e * qstride + j] = a == b ? S[a * N + a] : S[a * N + b] + S[b * N + a];
        }
    }
}

// _steer(..., force_arrive=True) (planner.py:296, 399-410) for one (node, target): the reference bounds
// it by wall-clock (:402-406); here by max_steps.  Sequential by nature: one thread.
template <typename T, class Sys, class Lqr>
__global__ void steer_force_kernel(Model<T, Sys, Lqr> md, const T* __restrict__ tstate, const T* __restrict__ tK,
                                   int cap, int id, const T* __restrict__ xtar, int max_steps,
                                   T* __restrict__ xs, T* __restrict__ us, int* __restrict__ len_out) {
    constexpr int N = Sys::N, M = Sys::M;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T x[N], xt[N], K[M * N];
    for (int d = 0; d < N; d++) { x[d] = tstate[(size_t)d * cap + id]; xt[d] = xtar[d]; }
    for (int k = 0; k < M * N; k++) K[k] = tK[(size_t)k * cap + id];
    int len = 0, steps = 0;
    while (true) {
        T e[N], u[M], uw[M], xn[N];
        erf_eval<T, N>(md.angle_mask, xt, x, e);
        for (int r = 0; r < M; r++) {
            T a = T(0);
            for (int d = 0; d < N; d++) a += K[r * N + d] * e[d];
            u[r] = a; uw[r] = a;
        }
        md.sys.step(x, uw, md.dt, xn);
        if (!feasible_eval<T, N, M>(md.feas, xn, u)) {
            len = fpr_keep(md, len);
            break;
        }
        steps += 1;
        if (steps > max_steps) break;
        bool close = true;                                        // np.allclose(x, xtar, 1e-4, 1e-4), planner.py:409
        for (int d = 0; d < N; d++) close = close && (abs_t(xn[d] - xt[d]) <= T(1e-4) + T(1e-4) * abs_t(xt[d]));
        if (close) break;
        for (int d = 0; d < N; d++) { x[d] = xn[d]; xs[len * N + d] = xn[d]; }
        for (int r = 0; r < M; r++) us[len * M + r] = u[r];
        len += 1;
        md.lqr.gain(md.sys, md.dt, x, u, K);
    }
    *len_out = len;
}

}  // namespace lq
