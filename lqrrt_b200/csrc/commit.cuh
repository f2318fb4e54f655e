// K4: Tree.add_node for a whole wave + the post-add bookkeeping
// (reference lqrrt/tree.py:77-96, lqrrt/planner.py:253-283), and the device side of the loop control
// (planner.py:286-311): sizing the next waves and stopping.
//
// Input: the exchange segments of all ranks (wave.cuh) -- per sample a meta word, and for samples that produced a
// node its parent and record row -- as written by the rollout kernels of the G replicas.  Every replica
// runs the same commit over the same segments: the nodes of a wave are appended in global sample order
// (segment = rank order, then sample order), so all replicas, and the 1-GPU run with the same wave sizes,
// hold bit-identical trees.  Node ids come from a prefix count over the meta words (no separate pack
// kernel, no atomics: slot = accepted samples with a lower global index).  The committing GPU copies the
// edges of ITS OWN segment from its rollout scratch into its edge slab and records the owner of every
// node, so edges are fetched from their owner when a plan is read out (tree.py:121-132) instead of
// crossing NVLink for every accepted node.
#pragma once
#include "steer.cuh"
#include "grid.cuh"

namespace lq {

template <typename T>
__device__ __forceinline__ void mark_path(const TreeView<T>& tv, int copy, int node) {
    // planner.py:266-270: ignores |= climb(node).  Stops at the first ancestor that is already marked:
    // whoever marked it also walks (or walked) on to the seed.
    while (node >= 0) {
        unsigned* word = (unsigned*)(tv.ignore[copy] + (node & ~3));
        const unsigned bit = 1u << (8 * (node & 3));
        const unsigned old = atomicOr(word, bit);
        if (old & bit) break;
        atomicAdd(tv.n_ignored2[copy], 1);
        node = tv.parent[node];
    }
}

// Stepwise driver: the host sizes the wave itself (lq_wave_expand).
static __global__ void ctl_set_kernel(WaveCtl* ctl, int slot, int W, int first, int count, int n_old) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->ring[slot].W = W; ctl->ring[slot].first = first; ctl->ring[slot].count = count; ctl->ring[slot].n_old = n_old;
    ctl->stopped = 0;
}

// Device-scheduled driver, plan start: waves 0 and 1 are sized from the tree as loaded (planner.py:172).
static __global__ void ctl_start_kernel(WaveCtl* ctl, const int* d_size, WavePolicy pol) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = *d_size;
    const int W = policy_wave(pol, n);
    for (int k = 0; k < CTL_RING; k++) {
        WaveSlot s{0, 0, 0, n};
        if (k < 2) { s.W = W; policy_slice(W, pol.rank, pol.world, &s.first, &s.count); }
        ctl->ring[k] = s;
    }
    ctl->stopped = 0; ctl->acc[0] = ctl->acc[1] = ctl->acc[2] = 0; ctl->done_steer[0] = ctl->done_steer[1] = 0; ctl->done_commit = 0;
}

constexpr int COMMIT_BLOCK = 128;
constexpr int COMMIT_SLOTS = 32;                          // sample slots per CTA: its 4 warps share their accepted records
constexpr unsigned PEER_SPIN_LIMIT_NS = 4000000000u;      // 4 s: a peer that never arrives is an error, not a hang

// One CTA per 32 consecutive global sample slots.  Every warp derives the node ids of those slots (prefix count of
// the accepted samples before them: one pass over the meta words, 4 KB per 1024 samples, L2 resident); the records
// are then dealt out to the CTA's four warps, all 32 lanes of a warp copy one record.  The last CTA to finish
// publishes the new tree size, the wave's counters and the control slots of the waves to come (every CTA has
// read the old size by then).
template <typename T>
__global__ void __launch_bounds__(COMMIT_BLOCK)
commit_x_kernel(TreeView<T> tv, WaveCtl* __restrict__ ctl, int wave_k,
                const char* segs, size_t seg_stride, int n_seg, int my_seg,
                const T* __restrict__ xseq, const T* __restrict__ useq, int Wc, int N, int M, int H,
                const T* __restrict__ goal_lo, const T* __restrict__ goal_hi,
                WavePolicy pol, DevStats* __restrict__ out, unsigned out_seq,
                const unsigned* __restrict__ arrive, unsigned seq,
                int* __restrict__ horizon, int hfactor, int hlo, int hhi, int copy, int lag2) {
    __shared__ int s_part[COMMIT_BLOCK / 32];
    __shared__ int s_err;
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();                          // launched early (PDL): the rollout kernel in front has finished from here on
    if (threadIdx.x == 0) s_err = 0;
    if (arrive) {                        // fused exchange: wait until every peer's results of this wave landed
        __syncthreads();
        if (threadIdx.x < n_seg) {
            const volatile unsigned* f = arrive + threadIdx.x;
            unsigned long long t0 = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while ((int)(*f - seq) < 0) {
                __nanosleep(32);
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > (unsigned long long)PEER_SPIN_LIMIT_NS) { s_err = 1; break; }
            }
        }
        __syncthreads();
        __threadfence_system();
    }
    const int old_size = *tv.size;
    const int rw = rec_words<T>(N, M);
    // a plan that has stopped appends nothing more (with pipelined waves the rollouts of the wave after the stop were
    // already in flight when the stop was decided)
    const bool dead = pol.device_schedule && ctl->stopped != 0;
    if (lag2) {                          // replay the goal paths of the commit before into this commit's copy of the marks
        const int n = min(*tv.goal_count[copy ^ 1], tv.goal_cap);
        for (int i = blockIdx.x * COMMIT_BLOCK + threadIdx.x; i < n; i += gridDim.x * COMMIT_BLOCK)
            mark_path(tv, copy, tv.goal_list[copy ^ 1][i]);
    }

    // global slot g -> (segment, sample); segments hold HDR_SAMPLED samples each.  The CTA's threads split the
    // metas that precede its 32 slots; lane `lane` of every warp owns slot g0 + lane.
    const int g0 = blockIdx.x * COMMIT_SLOTS;
    const int g = g0 + lane;
    int seg = -1, idx = 0, off = 0, before = 0;
    unsigned meta = 0;
    for (int s = 0; s < n_seg; s++) {
        const char* base = segs + (size_t)s * seg_stride;
        const int c = min(max(((const int*)base)[HDR_SAMPLED], 0), Wc);
        const SegView<T> sv = seg_view<T>((void*)base, Wc, N, M);
        const int upto = min(c, g0 - off);                      // this segment's samples that precede the CTA's slots
        {                                                       // four meta words per load, several loads in flight
            const uint4* m4 = reinterpret_cast<const uint4*>(sv.meta);
            const int n4 = max(upto, 0) >> 2;
#pragma unroll 4
            for (int i = threadIdx.x; i < n4; i += COMMIT_BLOCK) {
                const uint4 v = m4[i];
                before += (meta_len(v.x) > 0) + (meta_len(v.y) > 0) + (meta_len(v.z) > 0) + (meta_len(v.w) > 0);
            }
            for (int i = (n4 << 2) + threadIdx.x; i < upto; i += COMMIT_BLOCK) before += meta_len(sv.meta[i]) > 0 ? 1 : 0;
        }
        if (g >= off && g < off + c) { seg = s; idx = g - off; meta = sv.meta[idx]; }
        off += c;
    }
    const int total_slots = off;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) s_part[warp] = before;
    __syncthreads();
    before = 0;
    for (int w = 0; w < COMMIT_BLOCK / 32; w++) before += s_part[w];
    const bool mine = seg >= 0 && meta_len(meta) > 0 && !dead;
    const unsigned acc_mask = __ballot_sync(0xffffffffu, mine);
    const int rank_in_cta = __popc(acc_mask & ((1u << lane) - 1u));
    const int node_mine = old_size + before + rank_in_cta;
    if (warp == 0) {                     // the CTA's contribution to the wave totals (once, not per warp)
        int tsum = seg >= 0 ? meta_tries(meta) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
        if (lane == 0) {
            if (tsum) atomicAdd(&ctl->acc[1], tsum);
            if (acc_mask) atomicAdd(&ctl->acc[0], __popc(acc_mask));
        }
    }

    // ---- append: record r of the CTA goes to warp r % 4; all 32 lanes copy one record
    unsigned todo = __ballot_sync(0xffffffffu, mine && node_mine < tv.limit && (rank_in_cta % (COMMIT_BLOCK / 32)) == warp);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int node = __shfl_sync(0xffffffffu, node_mine, src);
        const int sg = __shfl_sync(0xffffffffu, seg, src);
        const int k = __shfl_sync(0xffffffffu, idx, src);
        const int L = meta_len(__shfl_sync(0xffffffffu, meta, src));
        const SegView<T> sv = seg_view<T>((void*)(segs + (size_t)sg * seg_stride), Wc, N, M);
        const T* row = sv.rec + (size_t)k * rw;
        // issue every load of the record before the first dependent store: one round trip instead of one per chunk
        const int par = sv.parent[k];
        const int nrec = N + M * N;
        T rv[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { const int q = lane + 32 * c; rv[c] = q < nrec ? row[q] : T(0); }
        const bool own = sg == my_seg;     // my own rollout: its edge is still in this GPU's scratch
        const T* sx = xseq + (size_t)k * H * N; const T* su = useq + (size_t)k * H * M;
        const int nx = own ? L * N : 0, nu = own ? L * M : 0;
        T ex[4], eu[2];
#pragma unroll
        for (int c = 0; c < 4; c++) { const int q = lane + 32 * c; ex[c] = q < nx ? sx[q] : T(0); }
#pragma unroll
        for (int c = 0; c < 2; c++) { const int q = lane + 32 * c; eu[c] = q < nu ? su[q] : T(0); }
        const int pc = tv.cost[par];
        bool in_goal = true;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int q = lane + 32 * c;
            if (q < N) {
                tv.state[(size_t)q * tv.cap + node] = rv[c];
                in_goal = in_goal && (goal_lo[q] < rv[c]) && (rv[c] < goal_hi[q]);      // planner.py:442-447
            } else if (q < nrec) {
                tv.K[(size_t)(q - N) * tv.cap + node] = rv[c];
            }
        }
        for (int q = lane + 96; q < nrec; q += 32) tv.K[(size_t)(q - N) * tv.cap + node] = row[q];       // wide models
        in_goal = __all_sync(0xffffffffu, in_goal);
        // position for the spatial index (lanes dx, dy of the first chunk hold those dims)
        const float gpx = (float)__shfl_sync(0xffffffffu, rv[0], tv.grid.dx & 31);
        const float gpy = (float)__shfl_sync(0xffffffffu, rv[0], tv.grid.dy & 31);
        T* dx = tv.edge_x + (size_t)node * H * N; T* du = tv.edge_u + (size_t)node * H * M;
#pragma unroll
        for (int c = 0; c < 4; c++) { const int q = lane + 32 * c; if (q < nx) dx[q] = ex[c]; }
#pragma unroll
        for (int c = 0; c < 2; c++) { const int q = lane + 32 * c; if (q < nu) du[q] = eu[c]; }
        for (int q = lane + 128; q < nx; q += 32) dx[q] = sx[q];                                         // long edges
        for (int q = lane + 64; q < nu; q += 32) du[q] = su[q];
        if (tv.zstate) {                             // whitened copy: lane i computes z_i = sum_{j >= i} U_ij x_j, ascending j
            T a = T(0);                              // (the order and the fmas of nearest.cuh whiten_apply: same bits)
            for (int jj = 0; jj < N; jj++) {
                const T xj = __shfl_sync(0xffffffffu, rv[0], jj);
                if (lane < N && jj >= lane) a = fma(__ldg(tv.U + lane * N + jj), xj, a);
            }
            if (lane < N) tv.zstate[(size_t)lane * tv.cap + node] = a;
        }
        if (lane == 0) {
            const int c = pc + L;
            tv.parent[node] = par; tv.cost[node] = c; tv.edge_len[node] = L; tv.owner[node] = (uint8_t)sg;
            if (tv.grid.on) grid_insert(tv.grid, node, gpx, gpy);
            if (in_goal) {                                                  // planner.py:260-281
                atomicAdd(tv.n_goal, 1);
                atomicMin(tv.best, ((unsigned long long)(unsigned)c << 32) | (unsigned)node);
                mark_path(tv, copy, node);
                if (lag2) {
                    const int gi = atomicAdd(tv.goal_count[copy], 1);
                    if (gi < tv.goal_cap) tv.goal_list[copy][gi] = node; else atomicExch(&ctl->acc[2], 2);
                }
            }
        }
    }

    // ---- completion: the last CTA publishes
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (s_err) atomicExch(&ctl->acc[2], 1);
        s_last = atomicAdd(&ctl->done_commit, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const int total = *(volatile int*)&ctl->acc[0], tries = *(volatile int*)&ctl->acc[1];
        const int err = *(volatile int*)&ctl->acc[2];
        const int flags = ((const int*)segs)[HDR_FLAGS];          // rank 0's stop request rides in segment 0
        const int nsz = min(old_size + total, tv.limit);
        const unsigned long long b = *(volatile unsigned long long*)tv.best;
        int hz = 0;
        if (hfactor) {                 // planner.py:421-424: the wave's votes, in global sample order
            hz = *horizon;
            for (int s = 0; s < n_seg; s++) {
                const char* base = segs + (size_t)s * seg_stride;
                const SegView<T> sv = seg_view<T>((void*)base, Wc, N, M);
                const int ns = min(max(sv.hdr[HDR_SAMPLED], 0), Wc);
                for (int q = 0; q < ns; q++) hz = horizon_apply(hz, meta_vote(sv.meta[q]), hfactor, hlo, hhi);
            }
            *horizon = hz;
        }
        // loop control (planner.py:289-311): a stop request or a full tree ends the plan; otherwise size the
        // wave after next from the size the next wave starts with
        int stopped = ctl->stopped;
        if (flags) stopped = flags;
        else if (nsz >= tv.limit && !stopped) stopped = 1;
        if (pol.device_schedule) {
            WaveSlot nxt{0, 0, 0, nsz};
            if (!stopped) {
                nxt.W = policy_wave(pol, nsz);
                policy_slice(nxt.W, pol.rank, pol.world, &nxt.first, &nxt.count);
            } else {
                WaveSlot& s1 = ctl->ring[(wave_k + 1) % CTL_RING];    // the wave already sampling becomes a no-op too
                s1.W = 0; s1.count = 0;
            }
            ctl->ring[(wave_k + 2) % CTL_RING] = nxt;
            ctl->stopped = stopped;
        }
        if (lag2) *tv.goal_count[copy ^ 1] = 0;        // replayed by every CTA of this commit; the next commit refills it
        ctl->acc[0] = 0; ctl->acc[1] = 0; ctl->acc[2] = 0; ctl->done_commit = 0;
        __threadfence();
        *tv.size = nsz;
        // counters for the host (mapped host memory or a device block the host copies)
        out->v[0] = nsz; out->v[1] = total_slots; out->v[2] = nsz - old_size; out->v[3] = *(volatile int*)tv.n_goal;
        out->v[4] = b == ~0ull ? -1 : (int)(b & 0xffffffffu);
        out->v[5] = b == ~0ull ? -1 : (int)(b >> 32);
        out->v[6] = (old_size + total > tv.limit) ? 1 : 0;
        out->v[7] = tries;
        out->v[8] = pol.device_schedule ? stopped : flags;
        out->v[9] = hz;
        out->v[10] = err; out->v[11] = 0;
        __threadfence_system();
        *(volatile unsigned*)&out->seq = out_seq;
    }
}

}  // namespace lq
