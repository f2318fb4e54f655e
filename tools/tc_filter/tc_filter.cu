// Round-2 prototype, NOT part of the product build (nothing under lqrrt_b200/ includes it): a tensor-core lower
// bound for the nearest scan with exact re-ranking, as a self-contained test program.
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tc_filter tc_filter.cu && ./tc_filter
//
// Exact cost (csrc/nearest.cuh, float32 diagonal metric, pre-scaled operands):
//     c(q, v) = sum_d e_d^2,  e_d = q_d - v_d,  wrapped dims: e_d = min(|e_d|, P_d - |e_d|)
// Bound (numerics in tests/studies/tf32_bound_study.py): with features
//     F(x) = [ x_d - centre_d  (unwrapped dims) ...,  w cos(x_h), w sin(x_h)  (wrapped dims) ]
//     LB(q, v) = |F(q)|^2 + |F(v)|^2 - 2 F(q).F(v)  <=  c(q, v)      (chord: 2 - 2 cos(dh) <= dh^2)
// F(q).F(v) is a K = 8 GEMM: one mma.sync m16n8k8 TF32 per 16 queries x 8 nodes, evaluated as 3xTF32
// (hi*hi + hi*lo + lo*hi, error <= 2^-20.7 (|F(q)|^2 + |F(v)|^2) measured), made safe with kappa = 2^-19.
// A pair is re-evaluated exactly only if LB_safe < the query's running best; the result is the exact scan's
// (cost, lowest index) bit for bit -- checked below against a plain exact kernel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int N = 6;              // states
constexpr int HD = 2;             // wrapped dim (heading)
constexpr int KF = 8;             // features: 5 unwrapped + cos + sin, padded to 8
constexpr int WARPS = 4;
constexpr int G = 4;              // query groups of 16 per warp
constexpr int QW = 16 * G;        // queries per warp
constexpr int QB = QW * WARPS;    // queries per CTA
constexpr int TN = 128;           // nodes per staged tile
constexpr float KAPPA = 1.9073486328125e-06f;   // 2^-19
constexpr float TWO_PI = 6.283185307179586f;

struct Metric { float w[N]; float period; float centre[N]; };

__device__ __forceinline__ float reduce_period(float a) { return a - TWO_PI * rintf(a * 0.15915494309189535f); }
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// pre-scaled row exactly as the scan sees it
__device__ __forceinline__ void scaled_row(const float* __restrict__ base, size_t stride, int i, const Metric& m, float (&o)[N]) {
#pragma unroll
    for (int d = 0; d < N; d++) {
        float a = base[(size_t)d * stride + i];
        if (d == HD) a = reduce_period(a);
        o[d] = __fmul_rn(a, m.w[d]);
    }
}
__device__ __forceinline__ float exact_cost(const float* q, const float* v, float period) {
    float c = 0.f;
#pragma unroll
    for (int d = 0; d < N; d++) {
        float e = q[d] - v[d];
        if (d == HD) { const float a = fabsf(e); e = fminf(a, period - a); }
        c = fmaf(e, e, c);
    }
    return c;
}
// features of a pre-scaled row: unwrapped dims centred, wrapped dim as w (cos, sin); |F|^2 in *n2
__device__ __forceinline__ void features(const float (&s)[N], const Metric& m, float (&f)[KF], float* n2) {
    int k = 0;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < N; d++) {
        if (d == HD) continue;
        f[k] = s[d] - m.centre[d] * m.w[d];
        acc = fmaf(f[k], f[k], acc);
        k++;
    }
    float sn, cs;
    sincosf(s[HD] / m.w[HD], &sn, &cs);          // s[HD] = w * heading (reduced to one period)
    f[k] = m.w[HD] * cs; acc = fmaf(f[k], f[k], acc); k++;
    f[k] = m.w[HD] * sn; acc = fmaf(f[k], f[k], acc); k++;
    for (; k < KF; k++) f[k] = 0.f;
    *n2 = acc;
}

// ---- reference: plain exact scan, one query per thread -----------------------------------------------
__global__ void exact_kernel(const float* __restrict__ state, int cap, int size, const float* __restrict__ queries, int W,
                             Metric m, float* __restrict__ cost, int* __restrict__ idx) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    float q[N];
    scaled_row(queries, W, j, m, q);
    float best = INFINITY; int b = -1;
    for (int i = 0; i < size; i++) {
        float v[N];
        scaled_row(state, cap, i, m, v);
        const float c = exact_cost(q, v, m.period);
        if (c < best) { best = c; b = i; }
    }
    cost[j] = best; idx[j] = b;
}

// ---- tensor-core bound + exact re-rank ------------------------------------------------------------------
// grid = nqb * nsplit CTAs; CTA (qb, split) scans nodes [split*chunk, min(+chunk, size)) for queries [qb*QB, +QB)
__global__ void __launch_bounds__(WARPS * 32)
filter_kernel(const float* __restrict__ state, int cap, int size, const float* __restrict__ queries, int W,
              Metric m, int nqb, int chunk, unsigned long long* __restrict__ partial /* [nsplit][W] */,
              unsigned long long* __restrict__ n_exact) {
    __shared__ __align__(16) float4 sB[TN / 8][32];          // per octet, per lane: {hi b0, hi b1, lo b0, lo b1}
    __shared__ __align__(8) float sNv[TN];                    // (1 - kappa) |F(v)|^2
    __shared__ float sV[TN][N];                               // pre-scaled node rows for the exact path
    __shared__ float sQ[QB][N];                               // pre-scaled query rows for the exact path
    __shared__ unsigned long long sBest[QB];                  // (cost bits << 32 | node), atomicMin

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int qb = blockIdx.x % nqb, split = blockIdx.x / nqb;
    const int n0 = split * chunk, n1 = min(n0 + chunk, size);
    const int q0 = qb * QB;

    for (int r = threadIdx.x; r < QB; r += blockDim.x) {
        float s[N];
        const int j = q0 + r;
        if (j < W) scaled_row(queries, W, j, m, s);
        else for (int d = 0; d < N; d++) s[d] = 0.f;
        for (int d = 0; d < N; d++) sQ[r][d] = s[d];
        sBest[r] = 0x7f800000ffffffffull;                     // (+inf, no node)
    }
    __syncthreads();

    // A fragments: rows = this warp's queries, cols = features; a0 (gid, tig) a1 (gid+8, tig) a2 (gid, tig+4) a3 (gid+8, tig+4)
    uint32_t Ah[G][4], Al[G][4];
    float nq[G][2];
#pragma unroll
    for (int g = 0; g < G; g++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = warp * QW + g * 16 + gid + 8 * h;
            float s[N], f[KF], n2;
#pragma unroll
            for (int d = 0; d < N; d++) s[d] = sQ[r][d];
            features(s, m, f, &n2);
            nq[g][h] = n2 * (1.f - KAPPA);
            const float x0 = f[tig], x1 = f[tig + 4];
            const uint32_t h0 = to_tf32(x0), h1 = to_tf32(x1);
            Ah[g][h] = h0; Ah[g][2 + h] = h1;
            Al[g][h] = to_tf32(x0 - __uint_as_float(h0)); Al[g][2 + h] = to_tf32(x1 - __uint_as_float(h1));
        }
    }

    unsigned long long local_exact = 0;
    for (int base = n0; base < n1; base += TN) {
        __syncthreads();
        {   // stage node `base + threadIdx.x`: exact row, features in fragment order, safe norm
            const int t = threadIdx.x, node = base + t;
            float s[N], f[KF], n2 = INFINITY;
            if (node < n1) { scaled_row(state, cap, node, m, s); features(s, m, f, &n2); }
            else { for (int d = 0; d < N; d++) s[d] = INFINITY; for (int k = 0; k < KF; k++) f[k] = 0.f; }
#pragma unroll
            for (int d = 0; d < N; d++) sV[t][d] = s[d];
            sNv[t] = node < n1 ? n2 * (1.f - KAPPA) : INFINITY;       // dead rows: bound = +inf, never pass
            float* dst = reinterpret_cast<float*>(&sB[t >> 3][0]);
#pragma unroll
            for (int k = 0; k < KF; k++) {                             // b0: (k = tig, n = gid), b1: (k = tig + 4, n = gid)
                const int l = (t & 7) * 4 + (k & 3);
                const uint32_t hi = to_tf32(f[k]);
                const uint32_t lo = to_tf32(f[k] - __uint_as_float(hi));
                dst[l * 4 + (k >> 2)] = __uint_as_float(hi);
                dst[l * 4 + 2 + (k >> 2)] = __uint_as_float(lo);
            }
        }
        __syncthreads();
        float bc[G][2];                                                // running best cost of my rows (may be stale: only costs passes)
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
            for (int h = 0; h < 2; h++)
                bc[g][h] = __uint_as_float((uint32_t)(sBest[warp * QW + g * 16 + gid + 8 * h] >> 32));

        const int octets = (min(TN, n1 - base) + 7) >> 3;
        for (int o = 0; o < octets; o++) {
            const float4 b = sB[o][lane];
            const float2 nv = *reinterpret_cast<const float2*>(&sNv[o * 8 + 2 * tig]);     // cols 2 tig, 2 tig + 1
            const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y), bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
#pragma unroll
            for (int g = 0; g < G; g++) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                mma_tf32(c, Al[g], bh0, bh1);
                mma_tf32(c, Ah[g], bl0, bl1);
                mma_tf32(c, Ah[g], bh0, bh1);
                // c0 (gid, 2 tig) c1 (gid, 2 tig + 1) c2 (gid + 8, 2 tig) c3 (gid + 8, 2 tig + 1)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int h = i >> 1;
                    const float lb = fmaf(-2.f, c[i], nq[g][h] + ((i & 1) ? nv.y : nv.x));
                    if (lb < bc[g][h]) {                               // rare: exact evaluation of this pair
                        const int r = warp * QW + g * 16 + gid + 8 * h;
                        const int tn = o * 8 + 2 * tig + (i & 1);
                        const float cst = exact_cost(sQ[r], sV[tn], m.period);
                        local_exact++;
                        if (cst < bc[g][h]) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(cst) << 32) | (uint32_t)(base + tn);
                            const unsigned long long old = atomicMin(&sBest[r], key);
                            bc[g][h] = fminf(cst, __uint_as_float((uint32_t)(old >> 32)));
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < QB; r += blockDim.x)
        if (q0 + r < W) partial[(size_t)split * W + q0 + r] = sBest[r];
    if (n_exact) atomicAdd(n_exact, local_exact);
}

__global__ void finalize_kernel(const unsigned long long* __restrict__ partial, int nsplit, int W, float* __restrict__ cost, int* __restrict__ idx) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    unsigned long long best = ~0ull;
    for (int s = 0; s < nsplit; s++) { const unsigned long long k = partial[(size_t)s * W + j]; if (k < best) best = k; }
    cost[j] = __uint_as_float((uint32_t)(best >> 32)); idx[j] = (int)(uint32_t)(best & 0xffffffffu);
}

int main(int argc, char** argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 8192, Nn = argc > 2 ? atoi(argv[2]) : 100000;
    const float arena = argc > 3 ? atof(argv[3]) : 200.f;
    printf("W = %d queries, N = %d nodes, n = %d, arena %.0f m\n", W, Nn, N, arena);
    std::vector<float> hs((size_t)N * Nn), hq((size_t)N * W);
    srand(1);
    auto U = [] { return (float)rand() / (float)RAND_MAX; };
    const float lo[N] = {0, 0, -7, 0.9f, -0.4f, -0.2f}, hi[N] = {arena, arena, 7, 1.1f, 0.4f, 0.2f};
    for (int i = 0; i < Nn; i++) for (int d = 0; d < N; d++) hs[(size_t)d * Nn + i] = lo[d] + (hi[d] - lo[d]) * U();
    for (int j = 0; j < W; j++) for (int d = 0; d < N; d++) hq[(size_t)d * W + j] = lo[d] + (hi[d] - lo[d]) * U();
    for (int i = 0; i < 64 && i < Nn; i++) for (int d = 0; d < N; d++) hs[(size_t)d * Nn + (Nn - 1 - i)] = hs[(size_t)d * Nn + i];   // exact ties
    Metric m;
    for (int d = 0; d < N; d++) { m.w[d] = 1.f; m.centre[d] = d == HD ? 0.f : 0.5f * (lo[d] + hi[d]); }
    m.period = TWO_PI * m.w[HD];
    float *ds, *dq, *c0, *c1; int *i0, *i1; unsigned long long *part, *nex;
    const int nqb = (W + QB - 1) / QB;
    int nsplit = (148 * 8 + nqb - 1) / nqb;
    int chunk = ((Nn + nsplit - 1) / nsplit + TN - 1) / TN * TN;
    nsplit = (Nn + chunk - 1) / chunk;
    CK(cudaMalloc(&ds, hs.size() * 4)); CK(cudaMalloc(&dq, hq.size() * 4));
    CK(cudaMalloc(&c0, W * 4)); CK(cudaMalloc(&c1, W * 4)); CK(cudaMalloc(&i0, W * 4)); CK(cudaMalloc(&i1, W * 4));
    CK(cudaMalloc(&part, (size_t)nsplit * W * 8)); CK(cudaMalloc(&nex, 8)); CK(cudaMemset(nex, 0, 8));
    CK(cudaMemcpy(ds, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dq, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms_exact = 0, ms_filter = 0;
    exact_kernel<<<(W + 127) / 128, 128>>>(ds, Nn, Nn, dq, W, m, c0, i0);
    CK(cudaEventRecord(e0)); exact_kernel<<<(W + 127) / 128, 128>>>(ds, Nn, Nn, dq, W, m, c0, i0); CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms_exact, e0, e1));
    filter_kernel<<<nqb * nsplit, WARPS * 32>>>(ds, Nn, Nn, dq, W, m, nqb, chunk, part, nex);
    finalize_kernel<<<(W + 127) / 128, 128>>>(part, nsplit, W, c1, i1);
    CK(cudaDeviceSynchronize());
    const int reps = 10;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) {
        filter_kernel<<<nqb * nsplit, WARPS * 32>>>(ds, Nn, Nn, dq, W, m, nqb, chunk, part, nullptr);
        finalize_kernel<<<(W + 127) / 128, 128>>>(part, nsplit, W, c1, i1);
    }
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms_filter, e0, e1)); ms_filter /= reps;
    std::vector<float> hc0(W), hc1(W); std::vector<int> hi0(W), hi1(W); unsigned long long hn = 0;
    CK(cudaMemcpy(hc0.data(), c0, W * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc1.data(), c1, W * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hi0.data(), i0, W * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hi1.data(), i1, W * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hn, nex, 8, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int j = 0; j < W; j++)
        if (hi0[j] != hi1[j] || memcmp(&hc0[j], &hc1[j], 4) != 0) { if (bad < 5) printf("MISMATCH q %d: exact (%g, %d) filter (%g, %d)\n", j, hc0[j], hi0[j], hc1[j], hi1[j]); bad++; }
    const double pairs = (double)W * Nn;
    printf("%s: %d of %d queries differ\n", bad ? "FAIL" : "PASS (bit-identical cost and index)", bad, W);
    printf("exact re-evaluations: %.3g pairs = %.3g of all pairs (%.1f per query per split, %d splits)\n", (double)hn, hn / pairs, (double)hn / W / nsplit, nsplit);
    printf("plain exact scan (1 query/thread, no tiling) %.3f ms; filter + finalize %.3f ms = %.3g pairs/s (product scan: ~1.8e12)\n",
           ms_exact, ms_filter, pairs / ms_filter * 1e3);
    return bad != 0;
}
