// K3: the default sampler closure of Planner.update_plan (reference
// lqrrt/planner.py:201-211) for a whole wave.
//
// Throughput mode draws from Philox4x32-10 (Salmon et al., SC'11) keyed by
// (seed, wave, global sample index, try): counter-based, so any split of a wave
// across GPUs draws identical numbers (oracle twin: oracle/philox.py).
// Reference-stream mode keeps NumPy's MT19937 on the host and only tests
// feasibility here (candidates_* kernels).
#pragma once
#include "steer.cuh"

namespace lq {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <typename T, int N> struct SamplerParams {
    T center[N], span[N], bias[N], goal[N];
    int tries;
};

// uniform on the 2^-24 grid of [0,1): exact in float and double
template <typename T> __device__ __forceinline__ T u24(uint32_t r) { return (T)(r >> 8) * (T)(1.0 / 16777216.0); }

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// One sample per group of TL threads: lane t evaluates try t of the reference's retry loop
// (planner.py:203-211) -- its own Philox draws, its own feasibility test -- and a group ballot picks the
// first feasible try (or the last one), exactly what the sequential loop would return.  The latency of a
// wave's sampling is one try instead of up to `tries` of them.
template <typename T, class Sys, class Lqr, int TL, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
sample_kernel(Model<T, Sys, Lqr> md, SamplerParams<T, Sys::N> sp, uint64_t seed, uint32_t wave,
              const WaveSlot* __restrict__ ws, T* __restrict__ samples, int sstride, int* __restrict__ tries_used,
              int first_known, int count_known) {
    constexpr int N = Sys::N, M = Sys::M;
    const int j = (blockIdx.x * BLOCK + threadIdx.x) / TL;
    const int t = threadIdx.x % TL;
    // this rank's slice of the wave: device-resident, or -- when the host already knows the wave's size for sure and
    // samples it ahead of the commit that would write the slot -- passed in (count_known >= 0)
    const int first = count_known >= 0 ? first_known : ws->first, count = count_known >= 0 ? count_known : ws->count;
    if (j >= count) return;                                  // whole groups leave together
    const uint32_t g = (uint32_t)(first + j);
    T x[N], u0[M];
#pragma unroll
    for (int i = 0; i < M; i++) u0[i] = T(0);
    bool ok = false;
    if (t < sp.tries) {
        T uni[(N + 4) / 4 * 4];
#pragma unroll
        for (int b = 0; b < (N + 4) / 4; b++) {
            uint32_t r[4];
            philox4x32_10(g, wave, (uint32_t)t, (uint32_t)b, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
            for (int k = 0; k < 4; k++) uni[4 * b + k] = u24<T>(r[k]);
        }
        // xrand = centers + spans * (U - 0.5), planner.py:204 -- un-fused so it matches NumPy bit for bit
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = add_rn(sp.center[i], mul_rn(sp.span[i], uni[i] - T(0.5)));
        const T r = uni[N];                                  // one draw gates every dim, planner.py:205-207
#pragma unroll
        for (int i = 0; i < N; i++) if (sp.bias[i] > r) x[i] = sp.goal[i];
        ok = feasible_eval<T, N, M>(md.feas, x, u0);          // planner.py:208-209
    }
    const unsigned votes = cg::tiled_partition<TL>(cg::this_thread_block()).ballot(ok);
    const int pick = votes ? __ffs(votes) - 1 : sp.tries - 1;   // first feasible try, else the last try
    if (t == pick) {
#pragma unroll
        for (int i = 0; i < N; i++) samples[(size_t)i * sstride + j] = x[i];
        if (tries_used) tries_used[j] = votes ? pick + 1 : sp.tries;
    }
}

// Reference-stream mode, step 1: feasibility of every candidate try.
template <typename T, class Sys, class Lqr>
__global__ void candidates_feasible_kernel(Model<T, Sys, Lqr> md, const T* __restrict__ cand, int cstride, int count,
                                           uint8_t* __restrict__ ok) {
    constexpr int N = Sys::N, M = Sys::M;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    T x[N], u0[M];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = cand[(size_t)i * cstride + j];
#pragma unroll
    for (int i = 0; i < M; i++) u0[i] = T(0);
    ok[j] = feasible_eval<T, N, M>(md.feas, x, u0) ? 1 : 0;
}

// Step 2: walk the candidates in stream order exactly as the reference's retry
// loop would (planner.py:203-211): each sample takes the first feasible of its
// next <= tries candidates, or the last of them.  Serial by definition; W*tries
// byte reads by one thread.
template <typename T, int N>
__global__ void candidates_select_kernel(const T* __restrict__ cand, int cstride, const uint8_t* __restrict__ ok,
                                         int n_cand, int count, int tries, T* __restrict__ samples, int sstride,
                                         int* __restrict__ pick, int* __restrict__ consumed) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int pos = 0;
        for (int j = 0; j < count; j++) {
            int chosen = -1;
            for (int t = 0; t < tries && pos < n_cand; t++) {
                chosen = pos++;
                if (ok[chosen]) break;
            }
            pick[j] = chosen;
        }
        *consumed = pos;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < count; j += blockDim.x) {
        const int c = pick[j];
        for (int i = 0; i < N; i++) samples[(size_t)i * sstride + j] = c >= 0 ? cand[(size_t)i * cstride + c] : T(0);
    }
}

}  // namespace lq
