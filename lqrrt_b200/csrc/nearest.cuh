// K1: nearest-node search = Planner._costs_to_go + argmin/argsort selection
// (reference lqrrt/planner.py:239-247, 340-350, erf_v :585-588), for a whole
// wave of queries against the SoA tree in HBM.
//
// Brute force on the FP32/FP64 pipes (no tensor cores: the angle wrap is not a
// bilinear form and K = n <= 12 would waste the MMA tile).  Work item =
// (query block, node split); node tiles are staged from the dim-major tree
// into shared memory once per CTA and broadcast to all lanes, every thread
// keeps Q queries and their running (min, argmin) in registers, partial
// results per split are reduced by nearest_finalize_kernel.
//
// Ignore semantics (planner.py:240-245: first node in cost order that is not
// on a finished plan, plain argmin if every node is ignored): ignored nodes are
// staged as +inf so they can never win; the "all ignored" case is detected from
// the engine's ignored-node counter and handled by scanning with the mask off.
#pragma once
#include "models.cuh"

namespace lq {

// Diagonal metric in pre-scaled form: cost = sum_i (w_i * wrap_i(q_i - v_i))^2,
// w_i = sqrt(S_ii); for an angle dim the wrapped magnitude of the scaled
// difference is min(|d|, P_i - |d|) with P_i = 2 pi w_i once both operands are
// reduced to one period.
template <typename T, int N> struct DiagMetric {
    T w[N];
    T period[N];
    uint32_t angle_mask;
    int exact;            // 1: unscaled d*S*d with rint wrap (float64 parity mode)
    T sdiag[N];
};

template <int N> struct PadDim { static constexpr int V = (N + 3) / 4 * 4; };

// What one scan launch covers.  Node range and query count are DEVICE scalars, so a launch can be enqueued
// before the wave that determines them has committed (the host only supplies upper bounds for the grid):
//   * old-tree scan of wave k+1: nodes [0, n_old) -- the tree as of the start of wave k, immutable -- on a
//     side stream while wave k's rollouts and commit run;
//   * delta scan: nodes [n_old, size) after commit k, on the wave's own stream;
//   * stepwise / stand-alone calls: d_lo = nullptr, the whole tree.
// Every CTA derives its node chunk from the actual range, so the work is balanced whatever the host's bound
// was.  Partial minima go to slots [slot0, slot0 + nsplit) of the partials arrays; consumers reduce them with
// a lexicographic (cost, index) minimum, so the lowest node index wins exact ties (np.argmin) in any order.
struct ScanArgs {
    const int* d_lo;         // first node (nullptr = 0)
    const int* d_hi;         // one past the last node
    const int* d_count;      // number of queries (nullptr = W)
    const int* d_nignored;   // non-null: plain argmin when every node is ignored (planner.py:244-245) is decided
                             // here; null: ignored nodes never win and the consumer handles "none left"
    int W;                   // host-side query count (bound): the grid was planned for it
    int nqb, nsplit, slot0;
    int use_ignore;
};

struct ScanRange { int qb, split, W, n0, n1; bool mask_on; };

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running and blocks in pdl_wait() until the predecessor has
// finished and its writes are visible; pdl_trigger() in the predecessor lets that early start happen as soon as every
// one of its CTAs has got going.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool scan_setup(const ScanArgs& a, int tile, int queries_per_cta, ScanRange& r) {
    pdl_trigger();                      // the rollout kernel behind a delta scan may set itself up meanwhile
    r.qb = blockIdx.x % a.nqb;
    r.split = blockIdx.x / a.nqb;
    r.W = a.d_count ? min(*a.d_count, a.W) : a.W;
    if (r.qb * queries_per_cta >= r.W) return false;             // no query in this block: nothing to write
    const int hi = *a.d_hi;
    const int lo = a.d_lo ? min(*a.d_lo, hi) : 0;
    int chunk = (hi - lo + a.nsplit - 1) / a.nsplit;
    chunk = (chunk + tile - 1) / tile * tile;
    r.n0 = min(lo + r.split * chunk, hi);
    r.n1 = min(r.n0 + chunk, hi);
    r.mask_on = a.use_ignore && (!a.d_nignored || *a.d_nignored < hi);
    return true;
}

template <typename T>
__device__ __forceinline__ T reduce_period(T a) {   // a -> a - 2 pi rint(a / 2 pi)
    return a - Num<T>::two_pi() * rint_t(a * Num<T>::inv_two_pi());
}

// Pre-scaling by sqrt(S_ii) is a rounded multiply of its own (never contracted into the later subtraction as
// an FMA): every float32 variant of the scan then evaluates round(q w) - round(v w) and they stay bit-identical
// to each other for any weights, not only for the demos' S = I.
__device__ __forceinline__ float prescale(float a, float w) { return __fmul_rn(a, w); }
__device__ __forceinline__ double prescale(double a, double w) { return __dmul_rn(a, w); }

// The per-dim arithmetic of the diagonal scans, shared with the per-query rescan of the rollout kernel
// (steer.cuh rescan_nearest) so both produce the same bits.
template <typename T>
__device__ __forceinline__ T diag_term_exact(T c, T e, T sd) { return c + e * sd * e; }     // float64 parity form: d S d
template <typename T>
__device__ __forceinline__ T diag_wrap_scaled(T e, T period) { const T a = abs_t(e); return min_t(a, period - a); }

// MASK >= 0: the erf angle mask is a compile-time constant (the common masks are instantiated, see
// engine.cuh), so only the wrapped dims pay for the wrap and the others are one FADD + one FFMA.
// MASK < 0: generic fallback reading met.angle_mask at run time.  EXACT selects the float64 parity
// formula d*S*d with an rint wrap; otherwise operands are pre-scaled by sqrt(S_ii) while staging.
template <typename T, int N, int Q, int BLOCK, int MASK, bool EXACT>
__global__ void __launch_bounds__(BLOCK)
nearest_diag_kernel(const T* __restrict__ state, int cap, const uint8_t* __restrict__ ignore, ScanArgs sa,
                    const T* __restrict__ queries, int qstride, DiagMetric<T, N> met,
                    T* __restrict__ pcost, int* __restrict__ pidx, int pstride) {
    constexpr int NP = PadDim<N>::V;
    constexpr int TILE = BLOCK;
    __shared__ __align__(16) T tile[TILE * NP];
    const uint32_t mask = MASK >= 0 ? (uint32_t)MASK : met.angle_mask;

    ScanRange sr;
    if (!scan_setup(sa, TILE, BLOCK * Q, sr)) return;
    const int qb = sr.qb, split = sa.slot0 + sr.split, W = sr.W, n0 = sr.n0, n1 = sr.n1;
    const bool mask_on = sr.mask_on;

    T q[Q][N];
    T best[Q];
    int bidx[Q];
    const int qbase = qb * (BLOCK * Q) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < Q; k++) {
        const int j = qbase + k * BLOCK;
        best[k] = Num<T>::inf();
        bidx[k] = -1;
#pragma unroll
        for (int d = 0; d < N; d++) {
            T v = j < W ? queries[(size_t)d * qstride + j] : T(0);
            if (!EXACT) {
                if (mask & (1u << d)) v = reduce_period(v);
                v = prescale(v, met.w[d]);
            }
            q[k][d] = v;
        }
    }

    for (int base = n0; base < n1; base += TILE) {
        __syncthreads();
        {   // stage one node per thread: coalesced along the node index for every dim
            const int node = base + threadIdx.x;
            const bool live = node < n1 && !(mask_on && ignore[node]);
            T v[NP];
#pragma unroll
            for (int d = 0; d < N; d++) {
                T a = live ? state[(size_t)d * cap + node] : T(0);
                if (!EXACT) {
                    if (mask & (1u << d)) a = reduce_period(a);
                    a = prescale(a, met.w[d]);
                }
                v[d] = live ? a : Num<T>::inf();       // dead rows can never win (inf cost or NaN)
            }
#pragma unroll
            for (int d = N; d < NP; d++) v[d] = T(0);
#pragma unroll
            for (int d = 0; d < NP; d++) tile[threadIdx.x * NP + d] = v[d];
        }
        __syncthreads();
        const int cnt = min(TILE, n1 - base);
#pragma unroll 2
        for (int i = 0; i < cnt; i++) {
            T v[NP];
#pragma unroll
            for (int d = 0; d < NP; d++) v[d] = tile[i * NP + d];   // same address in every lane: broadcast
            const int node = base + i;
#pragma unroll
            for (int k = 0; k < Q; k++) {
                T c = T(0);
#pragma unroll
                for (int d = 0; d < N; d++) {
                    T e = q[k][d] - v[d];
                    if (EXACT) {
                        if (mask & (1u << d)) e = reduce_period(e);
                        c = diag_term_exact(c, e, met.sdiag[d]);
                    } else {
                        if (mask & (1u << d)) e = diag_wrap_scaled(e, met.period[d]);
                        c = fma(e, e, c);
                    }
                }
                if (c < best[k]) { best[k] = c; bidx[k] = node; }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < Q; k++) {
        const int j = qbase + k * BLOCK;
        if (j < W) {
            pcost[(size_t)split * pstride + j] = best[k];
            pidx[(size_t)split * pstride + j] = bidx[k];
        }
    }
}

// ---- float32 scan on the packed FP32x2 pipe --------------------------------------------------------
// Same arithmetic as nearest_diag_kernel<float, ..., EXACT = false>, bit for bit (every operation is a
// round-to-nearest FADD / FFMA / FMUL on the same operands in the same order), issued differently:
//   * two queries share one instruction: Blackwell's add.f32x2 / fma.rn.f32x2 (FADD2 / FFMA2) work on a
//     64-bit register pair, so the 2n operations of a (query, node) pair cost n issue slots; the node
//     tile is staged pre-duplicated (v, v) so a broadcast LDS.128 delivers two aligned operand pairs;
//   * wrapped dims stay scalar (|e| is a free operand modifier there);
//   * the running argmin is two-level: inside a sub-block of SB nodes only the minimum cost is tracked
//     (one FMNMX per pair instead of compare + two selects), the (cost, sub-block) record is updated once
//     per sub-block with a strict '<' (earliest sub-block wins ties), and the winning sub-block is
//     re-evaluated once at the end to recover the first node whose cost equals the minimum -- the same
//     lowest-index tie-break as the one-level scan.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

template <int N, int MASK, bool PACKED>
__device__ __forceinline__ float2 pair_cost(const float2 (&q)[N], const float2* __restrict__ v, uint32_t mask,
                                            const float (&period)[N]) {
    // v[d] = (-node_d, -node_d) (staged negated: q - v == q + (-v) exactly); q[d] = (query A_d, query B_d)
    float2 c = f2(0.f, 0.f);
    bool first = true;
#pragma unroll
    for (int d = 0; d < N; d++) {
        const float2 vd = v[d];
        if (mask & (1u << d)) {
            float ea = q[d].x + vd.x, eb = q[d].y + vd.x;
            const float aa = fabsf(ea), ab = fabsf(eb);
            ea = fminf(aa, period[d] - aa); eb = fminf(ab, period[d] - ab);
            c.x = fmaf(ea, ea, c.x); c.y = fmaf(eb, eb, c.y);      // fma(e, e, 0) == e * e exactly
        } else if (PACKED) {
            const float2 e = __fadd2_rn(q[d], vd);
            c = first ? __fmul2_rn(e, e) : __ffma2_rn(e, e, c);
        } else {                                        // same operations, one lane at a time
            const float ea = q[d].x + vd.x, eb = q[d].y + vd.x;
            c.x = fmaf(ea, ea, c.x); c.y = fmaf(eb, eb, c.y);
        }
        first = false;
    }
    return c;
}

template <int N, int Q, int BLOCK, int MASK, int SB, int MINB, bool PACKED>
__global__ void __launch_bounds__(BLOCK, MINB)
nearest_diag_pk_kernel(const float* __restrict__ state, int cap, const uint8_t* __restrict__ ignore, ScanArgs sa,
                       const float* __restrict__ queries, int qstride, DiagMetric<float, N> met,
                       float* __restrict__ pcost, int* __restrict__ pidx, int pstride) {
    static_assert(Q % 2 == 0 && BLOCK % SB == 0, "queries are processed in pairs, tiles in whole sub-blocks");
    constexpr int NE = (N + 1) / 2 * 2;      // dims per node row, even: one LDS.128 = two (v, v) pairs
    constexpr int TILE = BLOCK;
    constexpr int QP = Q / 2;
    __shared__ __align__(16) float2 tile[TILE * NE];
    const uint32_t mask = MASK >= 0 ? (uint32_t)MASK : met.angle_mask;

    ScanRange sr;
    if (!scan_setup(sa, TILE, BLOCK * Q, sr)) return;
    const int qb = sr.qb, split = sa.slot0 + sr.split, W = sr.W, n0 = sr.n0, n1 = sr.n1;
    const bool mask_on = sr.mask_on;

    float2 q[QP][N];
    float best[Q];
    int bsb[Q];
    const int qbase = qb * (BLOCK * Q) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < Q; k++) {
        const int j = qbase + k * BLOCK;
        best[k] = INFINITY;
        bsb[k] = -1;
#pragma unroll
        for (int d = 0; d < N; d++) {
            float v = j < W ? queries[(size_t)d * qstride + j] : 0.f;
            if (mask & (1u << d)) v = reduce_period(v);
            v = prescale(v, met.w[d]);
            if (k & 1) q[k / 2][d].y = v; else q[k / 2][d].x = v;
        }
    }

    auto staged = [&](int node, float (&v)[N]) -> bool {       // the node row exactly as the scan sees it
        const bool live = node < n1 && !(mask_on && ignore[node]);
#pragma unroll
        for (int d = 0; d < N; d++) {
            float a = live ? state[(size_t)d * cap + node] : 0.f;
            if (mask & (1u << d)) a = reduce_period(a);
            a = prescale(a, met.w[d]);
            v[d] = live ? a : INFINITY;
        }
        return live;
    };

    for (int base = n0; base < n1; base += TILE) {
        __syncthreads();
        {
            float v[N];
            staged(base + threadIdx.x, v);
#pragma unroll
            for (int d = 0; d < NE; d++) tile[threadIdx.x * NE + d] = d < N ? f2(-v[d], -v[d]) : f2(0.f, 0.f);
        }
        __syncthreads();
        const int cnt = min(TILE, n1 - base);                    // rows past cnt are staged as +inf
        for (int s = 0; s < cnt; s += SB) {
            float2 m[QP];
#pragma unroll
            for (int kp = 0; kp < QP; kp++) m[kp] = f2(INFINITY, INFINITY);
#pragma unroll
            for (int i = 0; i < SB; i++) {
                float2 v[NE];
#pragma unroll
                for (int d = 0; d < NE; d++) v[d] = tile[(s + i) * NE + d];    // same address in every lane: broadcast
#pragma unroll
                for (int kp = 0; kp < QP; kp++) {
                    const float2 c = pair_cost<N, MASK, PACKED>(q[kp], v, mask, met.period);
                    m[kp].x = fminf(m[kp].x, c.x); m[kp].y = fminf(m[kp].y, c.y);
                }
            }
#pragma unroll
            for (int kp = 0; kp < QP; kp++) {
                if (m[kp].x < best[2 * kp]) { best[2 * kp] = m[kp].x; bsb[2 * kp] = base + s; }
                if (m[kp].y < best[2 * kp + 1]) { best[2 * kp + 1] = m[kp].y; bsb[2 * kp + 1] = base + s; }
            }
        }
    }
    // resolve the node inside the winning sub-block: first row whose cost equals the minimum
#pragma unroll
    for (int k = 0; k < Q; k++) {
        const int j = qbase + k * BLOCK;
        int bidx = -1;
        if (bsb[k] >= 0) {
            for (int i = 0; i < SB && bidx < 0; i++) {
                float v[N];
                float2 v2[NE];
                staged(bsb[k] + i, v);
#pragma unroll
                for (int d = 0; d < NE; d++) v2[d] = d < N ? f2(-v[d], -v[d]) : f2(0.f, 0.f);
                const float2 c = pair_cost<N, MASK, PACKED>(q[k / 2], v2, mask, met.period);
                if (((k & 1) ? c.y : c.x) == best[k]) bidx = bsb[k] + i;
            }
        }
        if (j < W) {
            pcost[(size_t)split * pstride + j] = best[k];
            pidx[(size_t)split * pstride + j] = bidx;
        }
    }
}

// ---- float32 scan with a partial-distance filter (experimental, LQ_NEAREST_PREFIX) ---------------------
// The cost chain c = fma(e_d, e_d, c), d = 0..N-1, only grows: every term is >= 0 and rounding is monotone,
// so the value after the first PREFIX dims (the positions) is a rigorous lower bound of the full cost IN
// FLOATING POINT.  A pair whose prefix is not below the query's running best cannot win and skips the other
// N - PREFIX dims; a pair that passes continues the same chain, so costs, argmin and the lowest-index
// tie-break are bit-identical to nearest_diag_kernel.  With nodes in insertion order the expected number of
// passes per (query, chunk of m nodes) is O(log m) (tests/studies/filter_study.py: 0.7 % of pairs on a C4 tree), but
// a warp pays for the long path whenever one of its lanes takes it -- whether that wins is a measurement
// for round 2.  Leading dims must not be wrapped (the engine checks the mask).
template <int N, int Q, int BLOCK, int MASK, int PREFIX>
__global__ void __launch_bounds__(BLOCK)
nearest_diag_prefix_kernel(const float* __restrict__ state, int cap, const uint8_t* __restrict__ ignore, ScanArgs sa,
                           const float* __restrict__ queries, int qstride, DiagMetric<float, N> met,
                           float* __restrict__ pcost, int* __restrict__ pidx, int pstride) {
    static_assert(PREFIX >= 1 && PREFIX < N, "the prefix is a proper, non-empty part of the state");
    constexpr int NP = PadDim<N>::V;
    constexpr int TILE = BLOCK;
    __shared__ __align__(16) float tile[TILE * NP];
    const uint32_t mask = MASK >= 0 ? (uint32_t)MASK : met.angle_mask;

    ScanRange sr;
    if (!scan_setup(sa, TILE, BLOCK * Q, sr)) return;
    const int qb = sr.qb, split = sa.slot0 + sr.split, W = sr.W, n0 = sr.n0, n1 = sr.n1;
    const bool mask_on = sr.mask_on;

    float q[Q][N];
    float best[Q];
    int bidx[Q];
    const int qbase = qb * (BLOCK * Q) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < Q; k++) {
        const int j = qbase + k * BLOCK;
        best[k] = INFINITY;
        bidx[k] = -1;
#pragma unroll
        for (int d = 0; d < N; d++) {
            float v = j < W ? queries[(size_t)d * qstride + j] : 0.f;
            if (mask & (1u << d)) v = reduce_period(v);
            q[k][d] = prescale(v, met.w[d]);
        }
    }

    for (int base = n0; base < n1; base += TILE) {
        __syncthreads();
        {
            const int node = base + threadIdx.x;
            const bool live = node < n1 && !(mask_on && ignore[node]);
#pragma unroll
            for (int d = 0; d < NP; d++) {
                float a = (live && d < N) ? state[(size_t)d * cap + node] : 0.f;
                if (d < N && (mask & (1u << d))) a = reduce_period(a);
                if (d < N) a = prescale(a, met.w[d]);
                tile[threadIdx.x * NP + d] = d < N ? (live ? a : INFINITY) : 0.f;
            }
        }
        __syncthreads();
        const int cnt = min(TILE, n1 - base);
#pragma unroll 2
        for (int i = 0; i < cnt; i++) {
            float v[NP];
#pragma unroll
            for (int d = 0; d < NP; d++) v[d] = tile[i * NP + d];
#pragma unroll
            for (int k = 0; k < Q; k++) {
                float c = 0.f;
#pragma unroll
                for (int d = 0; d < PREFIX; d++) {
                    const float e = q[k][d] - v[d];
                    c = fmaf(e, e, c);
                }
                if (c < best[k]) {                          // otherwise the full cost is >= best as well
#pragma unroll
                    for (int d = PREFIX; d < N; d++) {
                        float e = q[k][d] - v[d];
                        if (mask & (1u << d)) {
                            const float a = fabsf(e);
                            e = fminf(a, met.period[d] - a);
                        }
                        c = fmaf(e, e, c);
                    }
                    if (c < best[k]) { best[k] = c; bidx[k] = base + i; }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < Q; k++) {
        const int j = qbase + k * BLOCK;
        if (j < W) {
            pcost[(size_t)split * pstride + j] = best[k];
            pidx[(size_t)split * pstride + j] = bidx[k];
        }
    }
}

// d' S d with S symmetrised and packed (diagonal, then doubled upper triangle); shared with steer.cuh rescan_nearest
template <typename T, int N>
__device__ __forceinline__ T dense_cost(const T (&e)[N], const T* S) {
    T c = T(0);
    int s = N;
#pragma unroll
    for (int a = 0; a < N; a++) {
        T r = S[a] * e[a];
#pragma unroll
        for (int b = a + 1; b < N; b++) r = fma(S[s++], e[b], r);
        c = fma(e[a], r, c);
    }
    return c;
}

// Dense per-query metric: cost = d' S d with S symmetrised and packed
// (diagonal, then doubled upper triangle), NS = N (N + 1) / 2 coefficients per
// query held in registers.
template <typename T, int N, int BLOCK, int MASK>
__global__ void __launch_bounds__(BLOCK)
nearest_dense_kernel(const T* __restrict__ state, int cap, const uint8_t* __restrict__ ignore, ScanArgs sa,
                     const T* __restrict__ queries, int qstride,
                     const T* __restrict__ Spacked, uint32_t angle_mask_rt,
                     T* __restrict__ pcost, int* __restrict__ pidx, int pstride) {
    constexpr int NP = PadDim<N>::V;
    constexpr int NS = N * (N + 1) / 2;
    constexpr int TILE = BLOCK;
    __shared__ __align__(16) T tile[TILE * NP];
    const uint32_t angle_mask = MASK >= 0 ? (uint32_t)MASK : angle_mask_rt;

    ScanRange sr;
    if (!scan_setup(sa, TILE, BLOCK, sr)) return;
    const int qb = sr.qb, split = sa.slot0 + sr.split, W = sr.W, n0 = sr.n0, n1 = sr.n1;
    const bool mask_on = sr.mask_on;
    const int j = qb * BLOCK + threadIdx.x;

    T q[N], S[NS];
#pragma unroll
    for (int d = 0; d < N; d++) q[d] = j < W ? queries[(size_t)d * qstride + j] : T(0);
#pragma unroll
    for (int s = 0; s < NS; s++) S[s] = j < W ? Spacked[(size_t)s * qstride + j] : T(0);
    T best = Num<T>::inf();
    int bidx = -1;

    for (int base = n0; base < n1; base += TILE) {
        __syncthreads();
        {
            const int node = base + threadIdx.x;
            const bool live = node < n1 && !(mask_on && ignore[node]);
#pragma unroll
            for (int d = 0; d < NP; d++)
                tile[threadIdx.x * NP + d] = d < N ? (live ? state[(size_t)d * cap + node] : Num<T>::inf()) : T(0);
        }
        __syncthreads();
        const int cnt = min(TILE, n1 - base);
        for (int i = 0; i < cnt; i++) {
            T e[N];
#pragma unroll
            for (int d = 0; d < N; d++) {
                e[d] = q[d] - tile[i * NP + d];
                if (angle_mask & (1u << d)) e[d] = reduce_period(e[d]);
            }
            const T c = dense_cost<T, N>(e, S);
            // a dead (inf) node yields inf or NaN; neither compares below best
            if (c < best) { best = c; bidx = base + i; }
        }
    }
    if (j < W) {
        pcost[(size_t)split * pstride + j] = best;
        pidx[(size_t)split * pstride + j] = bidx;
    }
}

// Reduce the per-split partials in ascending split order with a strict '<':
// lowest node index wins ties, as NumPy's argmin / first-in-argsort does.
// (cost, index) candidates merge by the lexicographic minimum: associative and commutative, so partials can be
// combined in any order and the lowest node index still wins exact cost ties.
template <typename T>
__device__ __forceinline__ void merge_candidate(T& best, int& b, T c, int i) {
    if (i >= 0 && (b < 0 || c < best || (c == best && i < b))) { best = c; b = i; }
}

template <typename T>
__global__ void nearest_finalize_kernel(const T* __restrict__ pcost, const int* __restrict__ pidx, int pstride,
                                        int nsplit, int W, int* __restrict__ idx, T* __restrict__ cost) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    T best = Num<T>::inf();
    int b = -1;
    for (int s = 0; s < nsplit; s++) merge_candidate(best, b, pcost[(size_t)s * pstride + j], pidx[(size_t)s * pstride + j]);
    idx[j] = b < 0 ? 0 : b;
    if (cost) cost[j] = best;
}

// Fold the partial minima of slots [slot_lo, slot_lo + nslots) into the last of them, so that the rollout kernel has a
// single old-tree candidate per query to merge (reading hundreds of slots in its prologue cost more than its rollouts
// at small waves).  Block = 32 queries x 8 slot phases: every warp reads one slot row for 32 consecutive queries.
template <typename T>
__global__ void __launch_bounds__(256)
reduce_partials_kernel(T* __restrict__ pcost, int* __restrict__ pidx, int pstride, int slot_lo, int nslots, int W,
                       const int* __restrict__ d_count) {
    __shared__ T sc[8][32];
    __shared__ int si[8][32];
    const int lane = threadIdx.x & 31, phase = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    const int count = d_count ? min(*d_count, W) : W;
    T best = Num<T>::inf();
    int b = -1;
    if (j < count) {
#pragma unroll 4
        for (int s = slot_lo + phase; s < slot_lo + nslots; s += 8)
            merge_candidate(best, b, pcost[(size_t)s * pstride + j], pidx[(size_t)s * pstride + j]);
    }
    sc[phase][lane] = best; si[phase][lane] = b;
    __syncthreads();
    if (phase == 0 && j < count) {
#pragma unroll
        for (int p = 1; p < 8; p++) merge_candidate(best, b, sc[p][lane], si[p][lane]);
        const int out = slot_lo + nslots - 1;
        pcost[(size_t)out * pstride + j] = best;
        pidx[(size_t)out * pstride + j] = b;
    }
}

// ---- constant dense metric through the diagonal scans ---------------------------------------------------------
// cost = e^T S e with S = U^T U (U upper triangular, from a host-side Cholesky factorisation) is |U e|^2 = |U q - U v|^2:
// the tree keeps z = U v next to v, a wave's queries are transformed once, and the scan is the identity-metric
// diagonal scan over z.  One function produces every z (commit, load, seed, queries), with explicit fmas in a fixed
// order, so equal inputs give equal bits wherever they are transformed.
template <typename T, int N>
__device__ __forceinline__ void whiten_apply(const T* __restrict__ U, const T (&x)[N], T (&z)[N]) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        T a = T(0);
#pragma unroll
        for (int j = i; j < N; j++) a = fma(__ldg(U + i * N + j), x[j], a);
        z[i] = a;
    }
}

// z rows of nodes [first, *d_size) (or first + count_ub when d_size is null)
template <typename T, int N>
__global__ void whiten_range_kernel(const T* __restrict__ state, T* __restrict__ zstate, const T* __restrict__ U, int cap,
                                    int first, const int* __restrict__ d_size, int count_ub) {
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    const int hi = d_size ? *d_size : first + count_ub;
    if (i >= hi) return;
    T x[N], z[N];
#pragma unroll
    for (int d = 0; d < N; d++) x[d] = state[(size_t)d * cap + i];
    whiten_apply<T, N>(U, x, z);
#pragma unroll
    for (int d = 0; d < N; d++) zstate[(size_t)d * cap + i] = z[d];
}

template <typename T, int N>
__global__ void whiten_queries_kernel(const T* __restrict__ q, int qstride, int W, const int* __restrict__ d_count,
                                      const T* __restrict__ U, T* __restrict__ zq, int zstride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (d_count ? min(*d_count, W) : W)) return;
    T x[N], z[N];
#pragma unroll
    for (int d = 0; d < N; d++) x[d] = q[(size_t)d * qstride + j];
    whiten_apply<T, N>(U, x, z);
#pragma unroll
    for (int d = 0; d < N; d++) zq[(size_t)d * zstride + j] = z[d];
}

// Per-query metric for state-dependent providers (planner.py:348: S = lqr(x, 0)[0]
// at the sample), packed for nearest_dense_kernel.
template <typename T, class Sys, class Lqr>
__global__ void metric_kernel(Sys sys, Lqr lqr, T dt, const T* __restrict__ queries, int qstride, int W,
                              const int* __restrict__ d_count, T* __restrict__ Spacked) {
    constexpr int N = Sys::N;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (d_count ? min(*d_count, W) : W)) return;
    T x[N], S[N * N], u0[Sys::M];
#pragma unroll
    for (int d = 0; d < N; d++) x[d] = queries[(size_t)d * qstride + j];
#pragma unroll
    for (int r = 0; r < Sys::M; r++) u0[r] = T(0);            // planner.py:348: lqr(x, zeros)
    lqr.metric(sys, dt, x, u0, S);
    int s = N;
    for (int a = 0; a < N; a++) {
        Spacked[(size_t)a * qstride + j] = S[a * N + a];
        for (int b = a + 1; b < N; b++) Spacked[(size_t)(s++) * qstride + j] = S[a * N + b] + S[b * N + a];
    }
}

// K5: closest-to-guide (planner.py:313-318), one query, full (possibly
// column-zeroed, hence asymmetric) S.  Single CTA; once per plan.
template <typename T, int N>
struct FullMetric { T S[N * N]; T x[N]; uint32_t angle_mask; };

// Closest-to-guide (planner.py:313-318) over the whole tree: CLOSEST_CTAS CTAs take interleaved node blocks, each
// leaves its lexicographic (cost, index) minimum in scratch, the last one to finish merges them.  `counter` must be
// zero on entry and is zero again on exit.
constexpr int CLOSEST_CTAS = 64, CLOSEST_BLOCK = 256;
template <typename T, int N>
__global__ void __launch_bounds__(CLOSEST_BLOCK)
closest_kernel(const T* __restrict__ state, int cap, const int* __restrict__ d_size,
               FullMetric<T, N> fm, int* __restrict__ out, T* __restrict__ pc, int* __restrict__ pi,
               unsigned* __restrict__ counter) {
    __shared__ T sc[CLOSEST_BLOCK];
    __shared__ int si[CLOSEST_BLOCK];
    __shared__ bool s_last;
    const int size = *d_size;
    T best = Num<T>::inf();
    int b = -1;
    for (int node = blockIdx.x * CLOSEST_BLOCK + threadIdx.x; node < size; node += gridDim.x * CLOSEST_BLOCK) {
        T e[N];
#pragma unroll
        for (int d = 0; d < N; d++) {
            e[d] = fm.x[d] - state[(size_t)d * cap + node];
            if (fm.angle_mask & (1u << d)) e[d] = reduce_period(e[d]);
        }
        T c = T(0);
#pragma unroll
        for (int k = 0; k < N; k++) {      // sum_k (sum_j e_j S_jk) e_k  (tensordot(diffs, S) * diffs)
            T r = T(0);
#pragma unroll
            for (int jj = 0; jj < N; jj++) r = fma(e[jj], fm.S[jj * N + k], r);
            c = fma(r, e[k], c);
        }
        if (b < 0 || c < best) { best = c; b = node; }
    }
    auto reduce = [&]() {
        sc[threadIdx.x] = best; si[threadIdx.x] = b;
        __syncthreads();
        for (int s = CLOSEST_BLOCK / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const T c2 = sc[threadIdx.x + s]; const int i2 = si[threadIdx.x + s];
                const T c1 = sc[threadIdx.x]; const int i1 = si[threadIdx.x];
                if (i2 >= 0 && (i1 < 0 || c2 < c1 || (c2 == c1 && i2 < i1))) { sc[threadIdx.x] = c2; si[threadIdx.x] = i2; }
            }
            __syncthreads();
        }
    };
    reduce();
    if (threadIdx.x == 0) {
        pc[blockIdx.x] = sc[0]; pi[blockIdx.x] = si[0];
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    best = Num<T>::inf(); b = -1;
    if (threadIdx.x < gridDim.x) { best = __ldcg(&pc[threadIdx.x]); b = __ldcg(&pi[threadIdx.x]); }
    __syncthreads();
    reduce();
    if (threadIdx.x == 0) { *out = si[0] < 0 ? 0 : si[0]; *counter = 0; }
}

}  // namespace lq
