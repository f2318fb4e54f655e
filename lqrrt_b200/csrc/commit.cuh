// K4: Tree.add_node for a whole wave + the post-add bookkeeping
// (reference lqrrt/tree.py:77-96, lqrrt/planner.py:253-283).
//
// pack:   compact the wave's accepted samples (edge length > 0), in sample
//         order, into a fixed-size record buffer -- the multi-GPU allgather
//         payload (every rank contributes one such segment per wave).  A record
//         is the node only (state, gain, parent, edge length: ~100 B at n=6);
//         the edge trajectory stays on the GPU that rolled it out.
// commit: append the records of all segments, segment order then record order
//         (= global sample order), to the SoA tree; goal test, ignore marks
//         along the new plan, best-plan tracking.  The committing GPU copies the
//         edges of ITS OWN segment from its rollout scratch into its edge slab
//         and records the owner of every node, so edges are fetched from their
//         owner when a plan is read out (tree.py:121-132) instead of crossing
//         NVLink for every accepted node.
#pragma once
#include "steer.cuh"

namespace lq {

constexpr int REC_HEADER_INTS = 32;
enum { HDR_COUNT = 0, HDR_SAMPLED = 1, HDR_TRIES = 2, HDR_FLAGS = 3 };

template <typename T> struct RecView {
    int* header;
    T* state;      // [N][Wc]
    T* K;          // [M*N][Wc]
    int* parent;   // [Wc]
    int* len;      // [Wc]
    uint8_t* votes;   // [Wc] per SAMPLE of the segment's slice (not per record): adaptive-horizon votes
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

template <typename T>
__host__ __device__ inline size_t rec_bytes(int Wc, int N, int M, int H) {
    size_t b = REC_HEADER_INTS * sizeof(int);
    b += align16(sizeof(T) * (size_t)N * Wc);
    b += align16(sizeof(T) * (size_t)M * N * Wc);
    b += align16(sizeof(int) * (size_t)Wc) * 2;
    b += align16((size_t)Wc);
    (void)H;
    return b;
}

template <typename T>
__host__ __device__ inline RecView<T> rec_view(void* base, int Wc, int N, int M, int H) {
    RecView<T> v;
    char* p = (char*)base;
    v.header = (int*)p; p += REC_HEADER_INTS * sizeof(int);
    v.state = (T*)p; p += align16(sizeof(T) * (size_t)N * Wc);
    v.K = (T*)p; p += align16(sizeof(T) * (size_t)M * N * Wc);
    v.parent = (int*)p; p += align16(sizeof(int) * (size_t)Wc);
    v.len = (int*)p; p += align16(sizeof(int) * (size_t)Wc);
    v.votes = (uint8_t*)p;
    (void)H;
    return v;
}

// pack: compact the wave's accepted samples (edge length > 0), in sample order.  One kernel: every CTA
// first counts the accepted samples that precede its own range (the length array is a few KB, L2
// resident), then scans its range and copies its records; the last CTA's running total is the segment's
// record count.  Deterministic (no atomics): slot = number of accepted samples with a lower index.
// Multi-GPU, fused exchange: instead of packing into a local segment that an NCCL all-gather then copies,
// the pack kernel stores every record straight into segment `rank` of EVERY peer's receive buffer over
// NVLink (peer-mapped symmetric memory), and the last CTA publishes the header and raises this rank's
// arrival flag on every peer.  The commit kernel of each GPU waits for all G flags of the wave.  One
// kernel does pack + exchange; the collective's launch and rendezvous latency leaves the wave.
constexpr int MAX_PEERS = 16;
struct PeerDst {
    char* seg[MAX_PEERS];          // peer p's receive segment for this rank, current parity
    unsigned* flag[MAX_PEERS];     // peer p's arrival flag for this rank, current parity
    int n;                         // 0: single destination `rv` (local segment), no flags
    unsigned seq;                  // wave sequence number written to the flags
    unsigned* done;                // local CTA-completion counter
};

template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
pack_kernel(const int* __restrict__ len, const int* __restrict__ tries_used, const int* __restrict__ ids,
            const T* __restrict__ xnew, const T* __restrict__ Knew, int ostride, int count,
            RecView<T> rv, int* __restrict__ sample_of_slot, int Wc, int N, int M, int H, int flags, PeerDst pd,
            const uint8_t* __restrict__ votes) {
    __shared__ int wsum[BLOCK / 32];
    __shared__ int wtri[BLOCK / 32];
    __shared__ int s_before, s_tries;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j0 = blockIdx.x * BLOCK;
    // accepted samples (and sampler tries) before this CTA's range
    int before = 0, tries = 0;
    for (int j = threadIdx.x; j < j0; j += BLOCK) { before += len[j] > 0; tries += tries_used[j]; }
    // ... plus this CTA's own element, for the block scan below
    const int j = j0 + threadIdx.x;
    const int mine = (j < count && len[j] > 0) ? 1 : 0;
    tries += j < count ? tries_used[j] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        tries += __shfl_xor_sync(0xffffffffu, tries, o);
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 0) { wsum[warp] = before; wtri[warp] = tries; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int b2 = 0, t2 = 0;
        for (int w = 0; w < BLOCK / 32; w++) { b2 += wsum[w]; t2 += wtri[w]; }
        s_before = b2; s_tries = t2;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;              // per-warp totals of this CTA's own flags
    __syncthreads();
    int woff = 0, total = 0;
    for (int w = 0; w < BLOCK / 32; w++) { if (w < warp) woff += wsum[w]; total += wsum[w]; }
    const int slot = s_before + woff + incl - mine;
    if (votes && j < count) {
        if (pd.n == 0) rv.votes[j] = votes[j];
        else for (int p = 0; p < pd.n; p++) rec_view<T>(pd.seg[p], Wc, N, M, H).votes[j] = votes[j];
    }
    if (pd.n == 0) {
        if (mine) {
            for (int d = 0; d < N; d++) rv.state[(size_t)d * Wc + slot] = xnew[(size_t)d * ostride + j];
            for (int k = 0; k < M * N; k++) rv.K[(size_t)k * Wc + slot] = Knew[(size_t)k * ostride + j];
            rv.parent[slot] = ids[j]; rv.len[slot] = len[j];
            sample_of_slot[slot] = j;
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
            rv.header[HDR_COUNT] = s_before + total; rv.header[HDR_SAMPLED] = count;
            rv.header[HDR_TRIES] = s_tries; rv.header[HDR_FLAGS] = flags;
        }
        return;
    }
    // ---- peer mode
    if (mine) {
        T st[LQ_MAX_N];
        for (int d = 0; d < N; d++) st[d] = xnew[(size_t)d * ostride + j];
        const int par = ids[j], L = len[j];
        sample_of_slot[slot] = j;
        for (int p = 0; p < pd.n; p++) {
            const RecView<T> pv = rec_view<T>(pd.seg[p], Wc, N, M, H);
            for (int d = 0; d < N; d++) pv.state[(size_t)d * Wc + slot] = st[d];
            for (int k = 0; k < M * N; k++) pv.K[(size_t)k * Wc + slot] = Knew[(size_t)k * ostride + j];
            pv.parent[slot] = par; pv.len[slot] = L;
        }
    }
    __shared__ bool last;
    // the grid's last CTA knows the totals; it parks them for whichever CTA finishes last
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { pd.done[1] = (unsigned)(s_before + total); pd.done[2] = (unsigned)s_tries; }
    __threadfence_system();                          // my stores are visible system-wide before I count myself done
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(pd.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        const int cnt = (int)((volatile unsigned*)pd.done)[1], tr = (int)((volatile unsigned*)pd.done)[2];
        for (int p = 0; p < pd.n; p++) {
            int* h = (int*)pd.seg[p];
            h[HDR_COUNT] = cnt; h[HDR_SAMPLED] = count; h[HDR_TRIES] = tr; h[HDR_FLAGS] = flags;
        }
        __threadfence_system();
        for (int p = 0; p < pd.n; p++) *(volatile unsigned*)pd.flag[p] = pd.seq;     // arrival, after the data
        *pd.done = 0;
    }
}

// A rank whose slice of the wave is empty still has to announce itself to its peers.
static __global__ void pack_empty_peer_kernel(PeerDst pd, int flags) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int p = 0; p < pd.n; p++) {
        int* h = (int*)pd.seg[p];
        h[HDR_COUNT] = 0; h[HDR_SAMPLED] = 0; h[HDR_TRIES] = 0; h[HDR_FLAGS] = flags;
    }
    __threadfence_system();
    for (int p = 0; p < pd.n; p++) *(volatile unsigned*)pd.flag[p] = pd.seq;
}

template <typename T> struct TreeView {
    T* state;          // [N][cap]
    T* K;              // [M*N][cap]
    int* parent;       // [cap]
    int* cost;         // [cap] cumulative recorded steps from the seed; T_plan = (1 + cost) dt
    int* edge_len;     // [cap]
    uint8_t* ignore;   // [cap] (word-aligned: marks are set with atomicOr on the containing word)
    uint8_t* owner;    // [cap] rank (record segment) whose GPU holds the node's edge
    T* edge_x;         // [cap][H*N]
    T* edge_u;         // [cap][H*M]
    int* size;         // device scalar
    int* n_ignored;    // device scalar
    int* n_goal;       // device scalar
    unsigned long long* best;   // (cost << 32 | node), ~0 = none
    unsigned* done;             // CTAs of the running commit that have finished (last one publishes)
    int cap;           // array stride (node_capacity rounded up to 4)
    int limit;         // node_capacity: appends beyond it are dropped and flagged
};

template <typename T>
__device__ __forceinline__ void mark_path(const TreeView<T>& tv, int node) {
    // planner.py:266-270: ignores |= climb(node).  Stops at the first ancestor that is already marked:
    // whoever marked it also walks (or walked) on to the seed.
    while (node >= 0) {
        unsigned* word = (unsigned*)(tv.ignore + (node & ~3));
        const unsigned bit = 1u << (8 * (node & 3));
        const unsigned old = atomicOr(word, bit);
        if (old & bit) break;
        atomicAdd(tv.n_ignored, 1);
        node = tv.parent[node];
    }
}

struct DevStats { int v[12]; };

// One warp per (segment, record); the last CTA to finish publishes the new tree size and the wave's
// counters (every CTA has read the old size by then).
template <typename T>
__global__ void commit_kernel(TreeView<T> tv, const char* __restrict__ records, size_t seg_bytes, int n_seg,
                              int my_seg, const int* __restrict__ sample_of_slot,
                              const T* __restrict__ xseq, const T* __restrict__ useq,
                              int Wc, int N, int M, int H,
                              const T* __restrict__ goal_lo, const T* __restrict__ goal_hi,
                              DevStats* __restrict__ out, const unsigned* __restrict__ arrive, unsigned seq,
                              int* __restrict__ horizon, int hfactor, int hlo, int hhi) {
    if (arrive) {                        // fused exchange: wait until every peer's records of this wave landed
        if (threadIdx.x < n_seg) {
            const volatile unsigned* f = arrive + threadIdx.x;
            while (*f < seq) __nanosleep(64);
        }
        __syncthreads();
        __threadfence_system();
    }
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int seg = w / Wc, k = w % Wc;
    const int old_size = *tv.size;
    int before = 0, mine = 0;
    if (seg < n_seg) {
        for (int s = 0; s <= seg; s++) {
            const int c = ((const int*)(records + (size_t)s * seg_bytes))[HDR_COUNT];
            if (s < seg) before += c; else mine = c;
        }
    }
    const int node = old_size + before + k;
    if (seg < n_seg && k < mine && node < tv.limit) {
        const RecView<T> rv = rec_view<T>((void*)(records + (size_t)seg * seg_bytes), Wc, N, M, H);
        const int L = rv.len[k], par = rv.parent[k];
        bool in_goal = true;
        for (int d = lane; d < N; d += 32) {
            const T v = rv.state[(size_t)d * Wc + k];
            tv.state[(size_t)d * tv.cap + node] = v;
            in_goal = in_goal && (goal_lo[d] < v) && (v < goal_hi[d]);      // planner.py:442-447
        }
        in_goal = __all_sync(0xffffffffu, in_goal);
        for (int q = lane; q < M * N; q += 32) tv.K[(size_t)q * tv.cap + node] = rv.K[(size_t)q * Wc + k];
        if (seg == my_seg) {               // my own rollout: its edge is still in this GPU's scratch
            const int j = sample_of_slot[k];
            const T* sx = xseq + (size_t)j * H * N; T* dx = tv.edge_x + (size_t)node * H * N;
            for (int q = lane; q < L * N; q += 32) dx[q] = sx[q];
            const T* su = useq + (size_t)j * H * M; T* du = tv.edge_u + (size_t)node * H * M;
            for (int q = lane; q < L * M; q += 32) du[q] = su[q];
        }
        if (lane == 0) {
            const int c = tv.cost[par] + L;
            tv.parent[node] = par; tv.cost[node] = c; tv.edge_len[node] = L; tv.owner[node] = (uint8_t)seg;
            if (in_goal) {                                                  // planner.py:260-281
                atomicAdd(tv.n_goal, 1);
                atomicMin(tv.best, ((unsigned long long)(unsigned)c << 32) | (unsigned)node);
                mark_path(tv, node);
            }
        }
    }
    // ---- completion: the last CTA publishes
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(tv.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        int total = 0, sampled = 0, tries = 0;
        for (int s = 0; s < n_seg; s++) {
            const int* h = (const int*)(records + (size_t)s * seg_bytes);
            total += h[HDR_COUNT]; sampled += h[HDR_SAMPLED]; tries += h[HDR_TRIES];
        }
        const int nsz = min(old_size + total, tv.limit);
        const unsigned long long b = *(volatile unsigned long long*)tv.best;
        out->v[0] = nsz; out->v[1] = sampled; out->v[2] = nsz - old_size; out->v[3] = *(volatile int*)tv.n_goal;
        out->v[4] = b == ~0ull ? -1 : (int)(b & 0xffffffffu);
        out->v[5] = b == ~0ull ? -1 : (int)(b >> 32);
        out->v[6] = (old_size + total > tv.limit) ? 1 : 0;
        out->v[7] = tries;
        out->v[8] = ((const int*)records)[HDR_FLAGS];
        int hz = 0;
        if (hfactor) {                 // planner.py:421-424: the wave's votes, in global sample order
            hz = *horizon;
            for (int s = 0; s < n_seg; s++) {
                const RecView<T> rv = rec_view<T>((void*)(records + (size_t)s * seg_bytes), Wc, N, M, H);
                const int ns = rv.header[HDR_SAMPLED];
                for (int q = 0; q < ns; q++) hz = horizon_apply(hz, rv.votes[q], hfactor, hlo, hhi);
            }
            *horizon = hz;
        }
        out->v[9] = hz;
        out->v[10] = out->v[11] = 0;
        *tv.done = 0;
        __threadfence();
        *tv.size = nsz;
    }
}

}  // namespace lq
