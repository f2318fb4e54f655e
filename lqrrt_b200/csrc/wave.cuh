// Shared layout of the wave pipeline: the device-resident control block that lets waves be enqueued
// ahead of the commits that size them, and the per-rank exchange segment that carries a wave's rollout
// results to every replica of the tree (reference loop: lqrrt/planner.py:233-283).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace lq {

constexpr int MAX_PEERS = 16;
constexpr int CTL_RING = 4;

// ring[k % CTL_RING] describes wave k.  It is written two commits ahead: commit k-2 knows the tree size
// N_{k-1} wave k-1 starts from, which is all wave k's sampler and old-tree scan need --
//   W      global wave size, W_k = policy(N_{k-1})            (0: the wave is a no-op)
//   first, count   this rank's contiguous slice of the wave   (waves.partition)
//   n_old  = N_{k-1}: the old-tree scan covers [0, n_old), the delta scan [n_old, N_k)
// -- so both can run on a side stream while wave k-1's rollouts and commit are still in flight.
struct WaveSlot { int W, first, count, n_old; };

struct WaveCtl {
    WaveSlot ring[CTL_RING];
    int stopped;                    // sticky: a commit saw a stop request (1 done, 2 kill) or filled the tree
    int acc[3];                     // per-commit accumulators: accepted, tries (reset by the publishing CTA)
    unsigned done_steer[2];         // CTA completion counters of the rollout kernels, by wave parity (two waves may be in
    unsigned done_commit;           // flight with pipelined waves); commits are serialised (reset by the last CTA)
};

// How the next waves are sized on the device: the host's waves.wave_size, evaluated in float64 like Python.
struct WavePolicy {
    int fixed;                      // > 0: every wave has this many samples
    int wmin, wmax;
    double growth;                  // W = clamp(int(growth * N), wmin, wmax)
    int world, rank;
    int device_schedule;            // 0: stepwise driver (the host sizes every wave itself)
};

__host__ __device__ inline int policy_wave(const WavePolicy& p, int tree_size) {
    if (p.fixed > 0) return p.fixed;
    double w = p.growth * (double)tree_size;
    if (w < (double)p.wmin) w = (double)p.wmin;
    if (w > (double)p.wmax) w = (double)p.wmax;
    return (int)w;
}
__host__ __device__ inline void policy_slice(int W, int rank, int world, int* first, int* count) {
    const int base = W / world, extra = W % world;            // waves.partition
    *count = base + (rank < extra ? 1 : 0);
    *first = rank * base + (rank < extra ? rank : extra);
}

// ---- exchange segment: one per rank per wave ---------------------------------------------------------
// Every rank's rollout kernel writes, for each sample i of its slice, one meta word (edge length, adaptive-
// horizon vote, sampler tries) and -- if the sample produced a node -- its parent and a record row
// (state, then the gain K row-major), straight into segment `rank` of EVERY replica's receive buffer.  Nothing is
// compacted before the exchange: the commit kernel of each replica derives the node ids from the meta words
// of all segments (rank order, then sample order = global sample order), so all replicas append the same
// nodes in the same order with no pack kernel and no collective call in between.
constexpr int SEG_HDR_INTS = 8;
enum { HDR_SAMPLED = 0, HDR_FLAGS = 1 };

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

__host__ __device__ inline unsigned meta_pack(int len, unsigned vote, int tries) {
    return ((unsigned)len & 0xffffu) | ((vote & 0xffu) << 16) | (((unsigned)tries & 0xffu) << 24);
}
__host__ __device__ inline int meta_len(unsigned m) { return (int)(m & 0xffffu); }
__host__ __device__ inline unsigned meta_vote(unsigned m) { return (m >> 16) & 0xffu; }
__host__ __device__ inline int meta_tries(unsigned m) { return (int)(m >> 24); }

template <typename T>
__host__ __device__ inline int rec_words(int N, int M) {         // row length in T words, a multiple of 16 bytes
    const int per16 = 16 / (int)sizeof(T);
    return (N + M * N + per16 - 1) / per16 * per16;
}

template <typename T>
__host__ __device__ inline size_t seg_bytes(int Wc, int N, int M) {
    return align16(SEG_HDR_INTS * sizeof(int)) + 2 * align16(sizeof(int) * (size_t)Wc) +
           align16(sizeof(T) * (size_t)Wc * rec_words<T>(N, M));
}

template <typename T> struct SegView {
    int* hdr;
    unsigned* meta;    // [Wc]
    int* parent;       // [Wc]
    T* rec;            // [Wc][rec_words]
};

template <typename T>
__host__ __device__ inline SegView<T> seg_view(void* base, int Wc, int N, int M) {
    SegView<T> v;
    char* p = (char*)base;
    v.hdr = (int*)p; p += align16(SEG_HDR_INTS * sizeof(int));
    v.meta = (unsigned*)p; p += align16(sizeof(int) * (size_t)Wc);
    v.parent = (int*)p; p += align16(sizeof(int) * (size_t)Wc);
    v.rec = (T*)p;
    (void)N; (void)M;
    return v;
}

// Where a rollout kernel sends its results: n == 1 is the engine's own segment (single GPU, or the send
// buffer of an NCCL all-gather); n == world with flags is the fused exchange over NVLink peer memory.
struct ExDst {
    char* seg[MAX_PEERS];
    unsigned* flag[MAX_PEERS];     // peer p's arrival flag for this rank (nullptr: no flags)
    int n;
    unsigned seq;                  // wave sequence number the flags are raised to
};

// Spatial index of the tree over two position-like state dims (csrc/grid.cuh): GRID_F x GRID_F fine cells, each the
// head of a singly linked list of the nodes inside it (`next`, by node id), occupancy bitmaps of the fine cells and of
// the GRID_C x GRID_C coarse cells (GRID_R x GRID_R fine cells each).  Border cells stand for everything beyond them.
// Nodes are linked in by whoever appends them (commit kernel, load / append paths); lists are only ever prepended to,
// with the link written before the head is published, so a search may walk them while a commit is appending.
constexpr int GRID_F = 256, GRID_C = 32, GRID_R = GRID_F / GRID_C;
struct GridView {
    int* head;          // [GRID_F * GRID_F], -1 = empty
    int* next;          // [cap]
    unsigned* fine;     // [GRID_F * GRID_F / 32] occupancy bits, cell = fy * GRID_F + fx
    unsigned* coarse;   // [GRID_C * GRID_C / 32]
    float x0, y0;       // origin of cell (0, 0)
    float inv_hx, inv_hy;   // cells per state unit
    float kx, ky;       // metric cost of one cell width squared: hx^2 S_xx, hy^2 S_yy
    int dx, dy;         // the two state dims
    int on;             // 0: no index is kept
};

template <typename T> struct TreeView {
    GridView grid;
    // Whitened copy of the states (nullptr: none): z = U x with S = U^T U for a metric S that does not depend on the
    // state, so that cost = |U (q - v)|^2 and the diagonal scans serve a constant dense metric (nearest.cuh whiten_apply)
    T* zstate;         // [N][cap]
    const T* U;        // [N][N] row-major, upper triangular
    T* state;          // [N][cap]
    T* K;              // [M*N][cap]
    int* parent;       // [cap]
    int* cost;         // [cap] cumulative recorded steps from the seed; T_plan = (1 + cost) dt
    int* edge_len;     // [cap]
    // Ignore marks (planner.py:270), [cap] bytes, word-aligned (set with atomicOr on the containing word).  Two copies:
    // with pipelined waves (lag 2) wave k searches the marks as of commit k-2 while commit k-1 is adding its own,
    // so commit j marks copy j & 1 -- its own goal paths directly, those of commit j-1 replayed from goal_list --
    // and wave k reads copy k & 1.  Everything else uses copy 0.
    uint8_t* ignore[2];
    int* n_ignored2[2];   // marks per copy
    int* goal_list[2];    // goal nodes appended by the last commit of each parity (for the replay), [goal_cap]
    int* goal_count[2];
    int goal_cap;
    uint8_t* owner;    // [cap] rank (record segment) whose GPU holds the node's edge
    T* edge_x;         // [cap][H*N]
    T* edge_u;         // [cap][H*M]
    int* size;         // device scalar
    int* n_goal;       // device scalar
    unsigned long long* best;   // (cost << 32 | node), ~0 = none
    int cap;           // array stride (node_capacity rounded up to 4)
    int limit;         // node_capacity: appends beyond it are dropped and flagged
};

// what the publishing CTA of a commit writes for the host: lq_wave_stats (12 ints) + sequence + padding
struct DevStats { int v[12]; unsigned seq; int pad[3]; };

}  // namespace lq
