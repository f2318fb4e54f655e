// K2: Planner._steer(ID, xtar, force_arrive=False) for a whole wave
// (reference lqrrt/planner.py:354-438; exact control flow SURVEY.md section 3.2),
// fused with the lqr provider, the dynamics step and is_feasible.
//
// One sample per group of LANES threads.  All lanes of a group carry the same
// state/gain in registers and execute the same instruction stream (results are
// bitwise identical across lanes); only the obstacle tables of is_feasible are
// split across the lanes, and the verdict is combined with a group ballot.
// LANES = 1 is thread-per-sample; LANES = 32 is warp-per-sample for heavy
// collision fields.  Rollout is latency/issue bound: tiny matrices, no tensor
// cores.
#pragma once
#include <cooperative_groups.h>
#include "models.cuh"

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
};

// K = lqr(x, u).K, cooperatively when the provider supports it (FD-Jacobian + Riccati) and the group has lanes.
template <typename T, class Sys, class Lqr, int LANES>
__device__ __forceinline__ void group_gain(const Model<T, Sys, Lqr>& md, const T* x, const T* u, T* K, T s, T c,
                                           T* coop, int lane) {
    if constexpr (Lqr::COOP && LANES > 1) {
        md.lqr.gain_coop(md.sys, md.dt, x, u, K, coop, lane, LANES, GroupSync<LANES>());
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

// Outputs per sample j (all optional except len):
//   xseq [j][H][N], useq [j][H][M]   recorded edge (sample-major)
//   len  [j]                          kept length after the FPR cut
//   xnew [N][ostride], Knew [M*N][ostride]   final state and K = lqr(xnew, u_last).K (planner.py:256-257)
// The node to steer from is either given (ids) or, inside a wave, reduced here from K1's per-split partial
// minima (pcost/pidx, ascending split order with a strict '<' so the lowest index wins ties, as
// nearest_finalize_kernel does) and published in ids_out for the pack stage.
template <typename T>
struct NearPartials { const T* pcost; const int* pidx; int pstride; int nsplit; int* ids_out; };

template <typename T, class Sys, class Lqr, int LANES, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
steer_kernel(Model<T, Sys, Lqr> md,
             const T* __restrict__ tstate, const T* __restrict__ tK, int cap,
             const int* __restrict__ ids, NearPartials<T> np, const T* __restrict__ targets, int tstride, int W,
             T* __restrict__ xseq, T* __restrict__ useq, int* __restrict__ len_out,
             T* __restrict__ xnew_out, T* __restrict__ Knew_out, int ostride, uint8_t* __restrict__ votes) {
    constexpr int N = Sys::N, M = Sys::M;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    T* coop = reinterpret_cast<T*>(coop_raw) + (size_t)(threadIdx.x / LANES) * Lqr::COOP_SCALARS;
    const int gid = (blockIdx.x * BLOCK + threadIdx.x) / LANES;
    const int lane = threadIdx.x % LANES;
    if (gid >= W) return;                  // whole groups leave together (BLOCK % LANES == 0)
    const int j = gid;
    int id;
    if (np.pidx) {
        T best = Num<T>::inf();
        int b = -1;
        for (int sp = 0; sp < np.nsplit; sp++) {
            const T c = np.pcost[(size_t)sp * np.pstride + j];
            const int i2 = np.pidx[(size_t)sp * np.pstride + j];
            if (i2 >= 0 && (b < 0 || c < best)) { best = c; b = i2; }
        }
        id = b < 0 ? 0 : b;
        if (lane == 0) np.ids_out[j] = id;
    } else {
        id = ids[j];
    }

    T x[N], xt[N], K[M * N];
#pragma unroll
    for (int d = 0; d < N; d++) { x[d] = tstate[(size_t)d * cap + id]; xt[d] = targets[(size_t)d * tstride + j]; }
#pragma unroll
    for (int k = 0; k < M * N; k++) K[k] = tK[(size_t)k * cap + id];

    // sin/cos of the heading-like dim are evaluated once per state and shared by erf, dynamics,
    // the hull rotation in is_feasible and the gain (they were ~half of a step's instructions)
    constexpr int HD = Sys::HEADING;
    T sh = T(0), ch = T(1), sg = T(0), cg = T(1);
    if (HD >= 0) { sincos_t(x[HD < 0 ? 0 : HD], &sh, &ch); sincos_t(xt[HD < 0 ? 0 : HD], &sg, &cg); }

    int len = 0, nrec = 0, i = 0;
    // every rollout of a wave starts from the same planner-level horizon (frozen at wave start)
    const bool adaptive = md.hfactor != 0;
    int hor = adaptive ? *md.horizon : md.H;
    unsigned vote = 0;
    T last_emag[N];
#pragma unroll
    for (int d = 0; d < N; d++) last_emag[d] = Num<T>::inf();
    T* xrow = xseq + (size_t)j * md.H * N;      // never NULL: the engine supplies scratch
    T* urow = useq + (size_t)j * md.H * M;

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
        const bool ok = group_all<LANES>(feasible_eval<T, N, M>(md.feas, xn, u, lane, LANES, HD, sn, cn));   // planner.py:393
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
        group_gain<T, Sys, Lqr, LANES>(md, x, u, K, sh, ch, coop, lane);   // planner.py:436
    }

    // The node is the last KEPT state and its gain is lqr(xnew, u_seq[-1]).K (planner.py:256-257).  Without an
    // FPR cut, x / K already hold exactly that; after a cut (0 < len < nrec) the kept tail is re-read from
    // the group's own edge rows (written by lane 0) and the gain recomputed by the whole group.
    if (xnew_out && len > 0 && len < nrec) {
        GroupSync<LANES>()();
        T uk[M];
#pragma unroll
        for (int d = 0; d < N; d++) x[d] = xrow[(len - 1) * N + d];
#pragma unroll
        for (int r = 0; r < M; r++) uk[r] = urow[(len - 1) * M + r];
        T s2 = T(0), c2 = T(1);
        if (HD >= 0) sincos_t(x[HD < 0 ? 0 : HD], &s2, &c2);
        group_gain<T, Sys, Lqr, LANES>(md, x, uk, K, s2, c2, coop, lane);
    }
    if (lane == 0) {
        len_out[j] = len;
        if (votes) votes[j] = (uint8_t)vote;
        if (xnew_out && len > 0) {
#pragma unroll
            for (int d = 0; d < N; d++) xnew_out[(size_t)d * ostride + j] = x[d];
#pragma unroll
            for (int k = 0; k < M * N; k++) Knew_out[(size_t)k * ostride + j] = K[k];
        }
    }
}

// Per-query metric S = lqr(x, 0).S for cooperative providers: LANES threads per query share one sweep
// (the thread-per-query metric_kernel in nearest.cuh serves the others).  Packed like metric_kernel.
template <typename T, class Sys, class Lqr, int LANES>
__global__ void __launch_bounds__(32)
metric_coop_kernel(Model<T, Sys, Lqr> md, const T* __restrict__ queries, int qstride, int W, T* __restrict__ Spacked) {
    constexpr int N = Sys::N, M = Sys::M;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    T* coop = reinterpret_cast<T*>(coop_raw) + (size_t)(threadIdx.x / LANES) * Lqr::COOP_SCALARS;
    const int j = (blockIdx.x * 32 + threadIdx.x) / LANES;
    const int lane = threadIdx.x % LANES;
    if (j >= W) return;
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
