// K1, indexed: the same nearest-node selection as the brute-force scans of nearest.cuh
// (reference lqrrt/planner.py:239-247, 340-350) -- the same per-pair arithmetic, the same lexicographic
// (cost, index) minimum -- evaluated only for the nodes that can win.
//
// For a diagonal metric the cost is a sum of non-negative terms accumulated in dim order, so the part that
// comes from two position-like dims (x, y) is a lower bound of the whole in floating point as well (every
// partial sum of the fma chain is <= the next).  The tree keeps a uniform grid over (x, y) (wave.cuh GridView):
// a query looks at a cell only if the smallest cost any point of the cell could have is below the best cost
// found so far.  Bounds are taken with a safety margin far above rounding (1e-3 of a cell, 1e-3 relative), a
// cell is skipped only when its bound is >= the best cost, and candidates merge lexicographically, so the
// result is the brute-force result bit for bit: lowest index among exactly equal costs, ignored nodes absent.
//
// One query per group of LANES threads:
//   pass 0   the occupied coarse cell with the smallest bound is searched first (a good first upper bound, also
//            for samples far outside the explored region);
//   pass 1   every other occupied coarse cell whose bound is below the best cost, found with group ballots; the
//            lanes split a coarse cell's fine rows, walk the linked node lists of the fine cells that survive
//            their own bound, and share the best candidate after every coarse cell.
// Latency bound by design (dependent loads along the lists); what it saves is the N in W x N.
#pragma once
#include <cooperative_groups.h>
#include "nearest.cuh"
#include "wave.cuh"

namespace lq {
namespace cg = cooperative_groups;

__device__ __forceinline__ float grid_coord(float v, float origin, float inv_h) {
    // position in cell units, clamped to a finite range (NaN -> 0)
    return fminf(fmaxf((v - origin) * inv_h, -1.0e6f), 1.0e6f);
}
__device__ __forceinline__ int grid_cell_of(float c) { return (int)fminf(fmaxf(c, 0.0f), (float)(GRID_F - 1)); }

// Link node `node` (position x, y) into the index: one exchange on the cell's head, then the link.  A wave drops
// hundreds of near-identical nodes into one cell (goal-biased samples), so there is no retry loop here; a search that
// reaches a node whose link is still GRID_PENDING (all links are reset to it with the tree) waits for the store that
// is already on its way.
constexpr int GRID_PENDING = (int)0xfefefefe;
__device__ __forceinline__ void grid_insert(const GridView& g, int node, float x, float y) {
    const int fx = grid_cell_of(grid_coord(x, g.x0, g.inv_hx)), fy = grid_cell_of(grid_coord(y, g.y0, g.inv_hy));
    const int cell = fy * GRID_F + fx;
    const int old = atomicExch(&g.head[cell], node);
    __stcg(&g.next[node], old);
    atomicOr(&g.fine[cell >> 5], 1u << (cell & 31));
    const int cc = (fy / GRID_R) * GRID_C + fx / GRID_R;
    atomicOr(&g.coarse[cc >> 5], 1u << (cc & 31));
}

template <typename T>
__global__ void grid_insert_range_kernel(TreeView<T> tv, int first, const int* __restrict__ d_size, int count_ub) {
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    const int hi = d_size ? *d_size : first + count_ub;
    if (i >= hi) return;
    const GridView& g = tv.grid;
    grid_insert(g, i, (float)tv.state[(size_t)g.dx * tv.cap + i], (float)tv.state[(size_t)g.dy * tv.cap + i]);
}

// distance (in cells, minus the margin) from coordinate c to the cells [lo, hi) of one axis; the border cells reach
// to infinity, as everything beyond the grid is filed under them
__device__ __forceinline__ float grid_axis_gap(float c, int lo, int hi, int ncell) {
    float d = 0.0f;
    if (lo > 0) d = fmaxf(d, (float)lo - c);
    if (hi < ncell) d = fmaxf(d, c - (float)hi);
    return fmaxf(d - 1.0e-3f, 0.0f);
}
__device__ __forceinline__ float grid_bound(float gx, float gy, float kx, float ky) {
    return (gx * gx * kx + gy * gy * ky) * 0.999f;
}

// query in the form the scans use (nearest.cuh): angle dims reduced to one period and every dim pre-scaled
// by sqrt(S_ii), or untouched in the float64 parity form
template <typename T, int N, bool EXACT>
__device__ __forceinline__ void diag_query(uint32_t mask, const DiagMetric<T, N>& met, const T (&xq)[N], T (&q)[N]) {
#pragma unroll
    for (int d = 0; d < N; d++) {
        T v = xq[d];
        if (!EXACT) { if (mask & (1u << d)) v = reduce_period(v); v = prescale(v, met.w[d]); }
        q[d] = v;
    }
}
// one (query, node) pair: the operations of nearest_diag_kernel / nearest_diag_pk_kernel, in their order
template <typename T, int N, bool EXACT>
__device__ __forceinline__ T diag_pair_cost(uint32_t mask, const DiagMetric<T, N>& met, const T (&q)[N],
                                            const T* __restrict__ tstate, int cap, int node) {
    T a[N];
#pragma unroll
    for (int d = 0; d < N; d++) a[d] = tstate[(size_t)d * cap + node];        // all loads in flight together
    T c = T(0);
#pragma unroll
    for (int d = 0; d < N; d++) {
        if (EXACT) {
            T e = q[d] - a[d];
            if (mask & (1u << d)) e = reduce_period(e);
            c = diag_term_exact(c, e, met.sdiag[d]);
        } else {
            T v = a[d];
            if (mask & (1u << d)) v = reduce_period(v);
            v = prescale(v, met.w[d]);
            T e = q[d] - v;
            if (mask & (1u << d)) e = diag_wrap_scaled(e, met.period[d]);
            c = fma(e, e, c);
        }
    }
    return c;
}

template <typename T, int N, int MASK, bool EXACT, int LANES, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
nearest_grid_kernel(const T* __restrict__ state, int cap, const uint8_t* __restrict__ ignore, ScanArgs sa,
                    const T* __restrict__ queries, int qstride, DiagMetric<T, N> met, GridView g,
                    T* __restrict__ pcost, int* __restrict__ pidx, int pstride) {
    static_assert(LANES == GRID_R, "one lane per fine row of a coarse cell");
    static_assert((GRID_C * GRID_C) % LANES == 0 && BLOCK % LANES == 0, "group layout");
    constexpr int FW = GRID_F * GRID_F / 32, CW = GRID_C * GRID_C / 32;
    __shared__ unsigned s_fine[FW];
    __shared__ unsigned s_coarse[CW];
    const uint32_t mask = MASK >= 0 ? (uint32_t)MASK : met.angle_mask;

    pdl_trigger();
    const int W = sa.d_count ? min(*sa.d_count, sa.W) : sa.W;
    const int j0 = blockIdx.x * (BLOCK / LANES);
    if (j0 >= W) return;
    for (int i = threadIdx.x; i < FW; i += BLOCK) s_fine[i] = g.fine[i];
    for (int i = threadIdx.x; i < CW; i += BLOCK) s_coarse[i] = g.coarse[i];
    __syncthreads();
    const int j = j0 + threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    if (j >= W) return;                                   // whole groups leave together
    const auto tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    const int hi = *sa.d_hi;
    const bool mask_on = sa.use_ignore && (!sa.d_nignored || *sa.d_nignored < hi);

    T xq[N], q[N];
#pragma unroll
    for (int d = 0; d < N; d++) xq[d] = queries[(size_t)d * qstride + j];
    diag_query<T, N, EXACT>(mask, met, xq, q);
    float px = 0.0f, py = 0.0f;                           // (no dynamic indexing of the register array)
#pragma unroll
    for (int d = 0; d < N; d++) { if (d == g.dx) px = (float)xq[d]; if (d == g.dy) py = (float)xq[d]; }
    const float cx = grid_coord(px, g.x0, g.inv_hx), cy = grid_coord(py, g.y0, g.inv_hy);

    T best = Num<T>::inf();
    int b = -1;
    auto share = [&]() {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            const T c2 = tile.shfl_xor(best, o);
            const int i2 = tile.shfl_xor(b, o);
            merge_candidate(best, b, c2, i2);
        }
    };
    auto coarse_bound = [&](int cc) {
        const int ccx = cc % GRID_C, ccy = cc / GRID_C;
        return grid_bound(grid_axis_gap(cx, ccx * GRID_R, (ccx + 1) * GRID_R, GRID_F),
                          grid_axis_gap(cy, ccy * GRID_R, (ccy + 1) * GRID_R, GRID_F), g.kx, g.ky);
    };
    // the group searches coarse cell cc: lane r takes fine row r
    auto search = [&](int cc) {
        const int ccx = cc % GRID_C, ccy = cc / GRID_C;
        const int fy = ccy * GRID_R + lane;
        const int bit0 = fy * GRID_F + ccx * GRID_R;                       // GRID_R consecutive bits of one word
        unsigned row = (s_fine[bit0 >> 5] >> (bit0 & 31)) & ((1u << GRID_R) - 1u);
        const float gy = grid_axis_gap(cy, fy, fy + 1, GRID_F);
        while (row) {
            const int i = __ffs(row) - 1;
            row &= row - 1;
            const int fx = ccx * GRID_R + i;
            const float lb = grid_bound(grid_axis_gap(cx, fx, fx + 1, GRID_F), gy, g.kx, g.ky);
            if (!((T)lb < best)) continue;                                 // nothing in this cell can win or tie
            // L2-coherent loads: a commit may be prepending to this list right now (its nodes are >= hi and skipped,
            // but their links must be read as written, not from a stale L1 line)
            int guard = cap;
            auto link = [&](int node) {
                int nx = __ldcg(&g.next[node]);
                for (int spin = 0; nx == GRID_PENDING && spin < (1 << 20); spin++) { __nanosleep(64); nx = __ldcg(&g.next[node]); }
                return nx == GRID_PENDING ? -1 : nx;
            };
            for (int node = __ldcg(&g.head[fy * GRID_F + fx]); node >= 0 && guard > 0; node = link(node), guard--) {
                if (node >= hi || (mask_on && ignore[node])) continue;     // appended after the snapshot / on a finished plan
                const T c = diag_pair_cost<T, N, EXACT>(mask, met, q, state, cap, node);
                // as the scans: a cost that is not below +inf never wins; the lowest index wins exact ties
                if (c < best || (c == best && node < b)) { best = c; b = node; }
            }
        }
    };

    // ---- pass 0: the nearest occupied coarse cell first
    float blb = 3.0e38f;
    int bc = -1;
    for (int cc = lane; cc < GRID_C * GRID_C; cc += LANES) {
        if (!((s_coarse[cc >> 5] >> (cc & 31)) & 1u)) continue;
        const float lb = coarse_bound(cc);
        if (lb < blb) { blb = lb; bc = cc; }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        const float l2 = tile.shfl_xor(blb, o);
        const int c2 = tile.shfl_xor(bc, o);
        if (c2 >= 0 && (bc < 0 || l2 < blb || (l2 == blb && c2 < bc))) { blb = l2; bc = c2; }
    }
    if (bc >= 0) { search(bc); share(); }

    // ---- pass 1: every other coarse cell that may hold a better node
    for (int t = 0; t < GRID_C * GRID_C / LANES; t++) {
        const int cc = t * LANES + lane;
        bool go = cc != bc && ((s_coarse[cc >> 5] >> (cc & 31)) & 1u);
        if (go) go = (T)coarse_bound(cc) < best;
        unsigned m = tile.ballot(go);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            search(t * LANES + src);
            share();
        }
    }
    if (lane == 0) {
        pcost[(size_t)sa.slot0 * pstride + j] = best;
        pidx[(size_t)sa.slot0 * pstride + j] = b;
    }
}

}  // namespace lq
