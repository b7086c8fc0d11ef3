// KNN two-pass path, second generation: the exact evaluation happens INSIDE the scan, while the
// refs are still in shared memory.
//
//   knn_scan_eval_kernel   the streaming filter of nbr_two_pass.cuh (a warp = 4 x 32 queries x one
//                          split of the refs). The (query, step) pairs the filter flags are not
//                          written to per-query lists any more: they are compacted with ballots
//                          into ONE small work queue per warp. At the end of every 128-ref tile the
//                          warp drains the queue: lane e takes item e -- whichever query it belongs
//                          to --, reads that query's constants from a per-warp table, re-reads the
//                          flagged groups from the tile that is still in the ring, evaluates them
//                          in the exact reference arithmetic and appends the candidates below the
//                          query's bound to the query's candidate list in global memory
//                          (slot from a shared-memory counter). Every lane of a drain round does
//                          useful work: the lists of different queries no longer have to be
//                          walked in lockstep, and nothing is gathered from L2.
//   knn_topk_kernel        one thread per query reads its candidate keys 16 at a time (layout
//                          [slot][query]: coalesced) and folds them into a sorted best-K in
//                          registers with the sorting networks of nbr_engine.cuh. Keys are
//                          sortable(distance) << 32 | index, so the result does not depend on
//                          the order in which the candidates were appended, and the lowest index
//                          wins ties.
//
// Queries with fewer than k candidates below an ESTIMATED bound, or with an overflowed list, are
// flagged for the exact redo kernels (knn.cu).
#pragma once
#include "nbr_engine.cuh"

namespace b200pci {

// candidate keys per (query, split): about 3x the expected count below the bound (k <= 4: ~8k with
// the guaranteed bound; k <= 16: ~30; k <= 32: ~52 with two splits)
__host__ __device__ constexpr int se_cand_cap(int K) { return K <= 4 ? 64 : (K <= 16 ? 96 : 160); }
constexpr int SE_QCAP = 256;     // work-queue entries per warp (a step adds at most 128)

struct ScanEvalParams {
    u64 *cand;           // [warps][cap][128]  candidate keys
    uint32_t *cand_cnt;  // [warps][128]       candidates appended (> cap: overflow)
    int cap;
};

// per-warp shared memory
template <int STAGES>
struct ScanEvalSmem {
    static constexpr size_t ring = (size_t)STAGES * 4 * NBR_TILE * sizeof(float);
    static constexpr size_t ctrl = 128;                        // mbarriers
    static constexpr size_t qtab = 5 * 128 * sizeof(float);    // fa, fb, fc, |q|^2, tau of 128 queries
    static constexpr size_t queue = SE_QCAP * sizeof(uint32_t);
    static constexpr size_t cnt = 128 * sizeof(uint32_t);
    static constexpr size_t total = ring + ctrl + qtab + queue + cnt;
};

// Drain the warp's work queue (qn items, all from the tile in ring stage `sX`): out of line, so
// that the scan loop keeps its registers.
template <int MODE>
__device__ __noinline__ void scan_eval_drain(const uint32_t *queue, int qn, const float *qtab,
                                             const float4 *sX, uint32_t tile_group0, int N,
                                             uint32_t *ccnt, u64 *cand_warp, uint32_t cap) {
    constexpr int G4 = NBR_TILE / 4;
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < qn; base += 32) {
        const int e = base + lane;
        const uint32_t item = (e < qn) ? queue[e] : 0u;
        uint32_t m8 = item & 0xffu;
        const uint32_t step = (item >> 8) & 3u, owner = item >> 10;
        QueryRegs q;
        q.fa = qtab[owner];
        q.fb = qtab[128 + owner];
        q.fc = qtab[256 + owner];
        q.s = qtab[384 + owner];
        const float tau = qtab[512 + owner];
        while (__any_sync(0xffffffffu, m8 != 0u)) {
            const bool has = m8 != 0u;
            const int bit = has ? (31 - __clz((int)m8)) : 0;  // highest bit = lowest group
            m8 &= ~(1u << bit);
            const uint32_t g = step * NBR_BLK + (uint32_t)(7 - bit);  // group inside the tile
            const float4 X = sX[g], Y = sX[G4 + g], Z = sX[2 * G4 + g];
            float d[4];
            const uint32_t i0 = (tile_group0 + g) * 4u;
            dist4<MODE>(q, X, Y, Z, i0, N, d);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (has && d[i] < tau) {
                    const uint32_t pos = atomicAdd(&ccnt[owner], 1u);
                    if (pos < cap) cand_warp[(size_t)pos * 128 + owner] = make_key(d[i], i0 + i);
                }
            }
        }
    }
}

template <int MODE, int STAGES>
__device__ __forceinline__ void nbr_scan_eval(const NbrParams &p, const ScanEvalParams &ep) {
    using SM = ScanEvalSmem<STAGES>;
    constexpr int QT = NBR_QT;
    constexpr int G4 = NBR_TILE / 4;  // float4 per row per stage
    extern __shared__ __align__(128) unsigned char smem[];  // one warp per CTA
    const int lane = threadIdx.x & 31;
    float *tiles = reinterpret_cast<float *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::ring);
    float *qtab = reinterpret_cast<float *>(smem + SM::ring + SM::ctrl);
    uint32_t *queue = reinterpret_cast<uint32_t *>(smem + SM::ring + SM::ctrl + SM::qtab);
    uint32_t *ccnt = queue + SE_QCAP;

    const int b = blockIdx.z, split = blockIdx.y;
    const size_t warp_linear = (size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int tile0 = split * p.tiles_per_split;
    const int ntiles = min(p.tiles_per_split, p.total_tiles - tile0);
    const float *ws = p.ws_ref + (size_t)b * 4 * p.Npad;
    constexpr uint32_t stage_bytes = 4 * NBR_TILE * sizeof(float);

    auto issue_tile = [&](int t) {  // lane 0
        const int s = t % STAGES;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            tma_load_1d(tiles + (size_t)(s * 4 + r) * NBR_TILE,
                        ws + (size_t)r * p.Npad + (size_t)(tile0 + t) * NBR_TILE,
                        NBR_TILE * sizeof(float), &full[s]);
    };
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        for (int t = 0; t < min(ntiles, STAGES); ++t) issue_tile(t);
    }

    QueryRegs q[QT];
    float thr[QT];
    const int qi0 = blockIdx.x * (QT * 32) + lane;  // slot j: + 32 * j
#pragma unroll
    for (int j = 0; j < QT; ++j) {
        const int qi = qi0 + 32 * j;
        float x = 0.f, y = 0.f, z = 0.f, t0 = __int_as_float(0xff800000);  // -inf: never flagged
        if (qi < p.S) {
            const float *src = p.q + b * p.q_sb + qi * p.q_sp;
            x = src[p.q_ox];
            y = src[p.q_oy];
            z = src[2 * p.q_sc];
            t0 = p.tau_in ? p.tau_in[(size_t)b * p.S + qi] : p.tau_uniform;
        }
        q[j].set(x, y, z, p.q_xzy != 0);
        thr[j] = q[j].threshold(t0);
        const int o = j * 32 + lane;  // owner id inside the warp
        qtab[o] = q[j].fa;
        qtab[128 + o] = q[j].fb;
        qtab[256 + o] = q[j].fc;
        qtab[384 + o] = q[j].s;
        qtab[512 + o] = t0;
        ccnt[o] = 0u;
    }
    __syncwarp();
    u64 *cand_warp = ep.cand + warp_linear * (size_t)ep.cap * 128;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int qn = 0;  // queue length (warp-uniform)

    constexpr int SPT = G4 / NBR_BLK;  // steps per tile
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (t / STAGES) & 1);
        const float4 *sX = reinterpret_cast<const float4 *>(tiles + (size_t)(s * 4) * NBR_TILE);
        float4 X = sX[0], Y = sX[G4], Z = sX[2 * G4], W = sX[3 * G4];
        const uint32_t tile_group0 = (uint32_t)(tile0 + t) * G4;
#pragma unroll 1
        for (int step = 0; step < SPT; ++step) {
            const float4 *gX = sX + step * NBR_BLK;
            uint32_t m8[QT];
#pragma unroll
            for (int j = 0; j < QT; ++j) m8[j] = 0u;
#pragma unroll
            for (int u = 0; u < NBR_BLK; ++u) {
                const float4 cX = X, cY = Y, cZ = Z, cW = W;
                // prefetch the next group (one group past the tile at the very end: harmless,
                // still inside this CTA's shared memory, never used)
                X = gX[u + 1];
                Y = gX[G4 + u + 1];
                Z = gX[2 * G4 + u + 1];
                W = gX[3 * G4 + u + 1];
#pragma unroll
                for (int j = 0; j < QT; ++j)
                    if (filter4(q[j], cX, cY, cZ, cW) < thr[j]) m8[j] |= (0x80u >> u);
            }
            // compact the flagged (query, step) pairs into the warp's queue
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                const bool f = m8[j] != 0u;
                const unsigned bal = __ballot_sync(0xffffffffu, f);
                if (f)
                    queue[qn + __popc(bal & lt_mask)] =
                        ((uint32_t)(j * 32 + lane) << 10) | ((uint32_t)step << 8) | m8[j];
                qn += __popc(bal);
            }
            if (qn > SE_QCAP - QT * 32) {  // the next step could overflow the queue
                __syncwarp();
                scan_eval_drain<MODE>(queue, qn, qtab, sX, tile_group0, p.N | (p.r_xzy ? NBR_N_XZY : 0), ccnt, cand_warp, (uint32_t)ep.cap);
                qn = 0;
                __syncwarp();
            }
        }
        // the tile leaves the ring: evaluate what it flagged, then refill the stage
        __syncwarp();
        if (qn > 0) {
            scan_eval_drain<MODE>(queue, qn, qtab, sX, tile_group0, p.N | (p.r_xzy ? NBR_N_XZY : 0), ccnt, cand_warp, (uint32_t)ep.cap);
            qn = 0;
        }
        __syncwarp();
        if (lane == 0 && t + STAGES < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile(t + STAGES);
        }
    }
    __syncwarp();
    uint32_t *pc = ep.cand_cnt + warp_linear * 128;
#pragma unroll
    for (int j = 0; j < QT; ++j) pc[j * 32 + lane] = ccnt[j * 32 + lane];
}

// ---- pass 2: top-k over the candidate lists (one thread per query) ------------------------------
struct TopkParams {
    void *idx;    // int64/int32 [B,S,kout]
    float *dist;  // nullable
    int idx_is_int64;
    int kout;
    int *fail_count;  // [0] flagged tiles, [1] flagged queries (pre-zeroed)
    int *fail_list;   // [B*S] redo flags (pre-zeroed), then the flagged-tile list, then the query list
    const u64 *cand;
    const uint32_t *cand_cnt;
    const int *qperm;  // [B][S] processed query row -> original row (null: identity)
    int scan_tiles;  // query tiles of the scan grid
    int nsplit;
    int cap;  // candidate list capacity per (query, split)
    int balance;  // top-k kernel: hand the queries of a CTA out in order of their list length
};

constexpr int TOPK_THREADS = 128;  // = the 128 queries of one scan warp
template <int K>
__global__ void __launch_bounds__(TOPK_THREADS) knn_topk_kernel(int S, TopkParams tp) {
    constexpr bool NET = K > 4;
    constexpr int NBLK = NET ? K / 16 : 1;
    constexpr int KR = NET ? 16 : K;
    static_assert(NBLK <= 2, "top-k kernel: K <= 32");
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.z, tile = blockIdx.x;
    // Which of the CTA's 128 queries this thread takes. A warp walks ceil(max n / 16) chunks, so 32
    // queries with list lengths drawn at random (mean ~61, maximum of 32 ~110 for k = 16) waste almost
    // half of the lanes' work: the queries are handed out in order of their total list length
    // (bitonic sort of 128 (length, query) words in shared memory), so a warp's 32 lists are alike.
    int tid = threadIdx.x;
    if (tp.balance) {
        __shared__ uint32_t order[TOPK_THREADS];
        uint32_t total = 0u;
        for (int s = 0; s < tp.nsplit; ++s) {
            const size_t warp_linear = (size_t)(b * tp.nsplit + s) * tp.scan_tiles + tile;
            total += min(tp.cand_cnt[warp_linear * 128 + tid], (uint32_t)tp.cap);
        }
        order[tid] = (total << 8) | (uint32_t)tid;
        __syncthreads();
        for (int k2 = 2; k2 <= TOPK_THREADS; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                if (threadIdx.x < TOPK_THREADS / 2) {
                    const int t = threadIdx.x;
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const uint32_t x = order[i], y = order[l];
                    if ((x > y) == ((i & k2) == 0)) {
                        order[i] = y;
                        order[l] = x;
                    }
                }
                __syncthreads();
            }
        }
        tid = (int)(order[threadIdx.x] & 0xffu);
    }
    const int qi = tile * TOPK_THREADS + tid;
    const bool valid = qi < S;
    u64 S0[KR], S1[NBLK > 1 ? 16 : 1];
#pragma unroll
    for (int i = 0; i < KR; ++i) S0[i] = B200PCI_KEY_INF;
    if constexpr (NBLK > 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) S1[i] = B200PCI_KEY_INF;
    }
    for (int s = 0; s < tp.nsplit; ++s) {
        const size_t warp_linear = (size_t)(b * tp.nsplit + s) * tp.scan_tiles + tile;
        const uint32_t ntot = tp.cand_cnt[warp_linear * 128 + tid];
        const int n = (int)min(ntot, (uint32_t)tp.cap);
        const u64 *col = tp.cand + warp_linear * (size_t)tp.cap * 128 + tid;
        if constexpr (NET) {
            const int nchunk = (warp_max_i(n) + 15) / 16;
            for (int c = 0; c < nchunk; ++c) {
                u64 C[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    C[i] = (c * 16 + i < n) ? col[(size_t)(c * 16 + i) * 128] : ~0ull;
                sort16(C);
                if constexpr (NBLK == 1) {
                    merge_low16(S0, C);
                } else {
                    merge_low16(S1, C);    // S1 = 16 smallest of (top block U chunk)
                    merge_full16(S0, S1);  // S0 = low half, S1 = high half
                }
            }
        } else {
            const int nmax = warp_max_i(n);
            for (int c = 0; c < nmax; ++c) {
                u64 key = (c < n) ? col[(size_t)c * 128] : ~0ull;
#pragma unroll
                for (int r = 0; r < K; ++r) ce64(S0[r], key);
            }
        }
    }
    const size_t qrow = (size_t)b * S + (valid ? qi : 0);
    const int kout = tp.kout;
    // rows flagged by knn_flag_kernel belong to the exact redo kernels (which may be running
    // concurrently on another stream): not written here
    const bool mine = valid && tp.fail_list[qrow] == 0;
    const size_t orow = tp.qperm ? (size_t)b * S + tp.qperm[qrow] : qrow;  // original query row
#pragma unroll
    for (int i = 0; i < K; ++i) {
        if (i < kout && mine) {
            u64 key;
            if constexpr (NBLK > 1)
                key = (i < 16) ? S0[i < 16 ? i : 0] : S1[i >= 16 ? i - 16 : 0];
            else
                key = S0[i];
            const size_t o = orow * kout + i;
            const uint32_t id = (uint32_t)key;
            if (tp.idx_is_int64)
                reinterpret_cast<long long *>(tp.idx)[o] = (long long)id;
            else
                reinterpret_cast<int *>(tp.idx)[o] = (int)id;
            if (tp.dist) tp.dist[o] = sortable2f((uint32_t)(key >> 32));
        }
    }
}

// The same for launches that leave most of the GPU idle (one frame pair: 128 CTAs of 4 warps, each
// thread walking the lists of all ref splits one after the other -- 53 us against 7 us per cloud in
// a 64-pair launch): one thread per (query, split group). blockDim = (128, P); thread (q, p) folds
// the lists of splits p, p + P, ...; groups p > 0 leave their sorted block(s) in shared memory and
// group 0 folds them in, 16 keys at a time (each block is sorted, which is all the fold needs).
// The selected set and its order do not depend on how the keys were grouped: same result.
// dynamic shared memory: (P - 1) * K * 128 keys.
template <int K>
__global__ void __launch_bounds__(TOPK_THREADS * 4) knn_topk_split_kernel(int S, TopkParams tp) {
    static_assert(K == 16 || K == 32, "split top-k: K = 16 or 32");
    constexpr int NBLK = K / 16;
    extern __shared__ __align__(16) u64 tk_lists[];  // [P-1][K][128]
    const int tid = threadIdx.x, part = threadIdx.y, P = blockDim.y;
    const int b = blockIdx.z, tile = blockIdx.x;
    const int qi = tile * TOPK_THREADS + tid;
    const bool valid = qi < S;
    u64 S0[16], S1[NBLK > 1 ? 16 : 1];
#pragma unroll
    for (int i = 0; i < 16; ++i) S0[i] = B200PCI_KEY_INF;
    if constexpr (NBLK > 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) S1[i] = B200PCI_KEY_INF;
    }
    auto fold = [&](u64(&C)[16]) {  // C sorted ascending
        if constexpr (NBLK == 1) {
            merge_low16(S0, C);
        } else {
            merge_low16(S1, C);    // S1 = 16 smallest of (top block U chunk)
            merge_full16(S0, S1);  // S0 = low half, S1 = high half
        }
    };
    for (int s = part; s < tp.nsplit; s += P) {
        const size_t warp_linear = (size_t)(b * tp.nsplit + s) * tp.scan_tiles + tile;
        const uint32_t ntot = tp.cand_cnt[warp_linear * 128 + tid];
        const int n = (int)min(ntot, (uint32_t)tp.cap);
        const u64 *col = tp.cand + warp_linear * (size_t)tp.cap * 128 + tid;
        const int nchunk = (warp_max_i(n) + 15) / 16;
        for (int c = 0; c < nchunk; ++c) {
            u64 C[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) C[i] = (c * 16 + i < n) ? col[(size_t)(c * 16 + i) * 128] : ~0ull;
            sort16(C);
            fold(C);
        }
    }
    if (part > 0) {
        u64 *mine = tk_lists + (size_t)(part - 1) * K * 128 + tid;
#pragma unroll
        for (int i = 0; i < 16; ++i) mine[(size_t)i * 128] = S0[i];
        if constexpr (NBLK > 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mine[(size_t)(16 + i) * 128] = S1[i];
        }
    }
    __syncthreads();
    if (part > 0) return;
    for (int o = 0; o + 1 < P; ++o) {
        const u64 *theirs = tk_lists + (size_t)o * K * 128 + tid;
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
            u64 C[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) C[i] = theirs[(size_t)(blk * 16 + i) * 128];
            fold(C);
        }
    }
    const size_t qrow = (size_t)b * S + (valid ? qi : 0);
    const int kout = tp.kout;
    const bool mine_row = valid && tp.fail_list[qrow] == 0;  // flagged rows belong to the exact redo
    const size_t orow = tp.qperm ? (size_t)b * S + tp.qperm[qrow] : qrow;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        if (i < kout && mine_row) {
            u64 key;
            if constexpr (NBLK > 1)
                key = (i < 16) ? S0[i < 16 ? i : 0] : S1[i >= 16 ? i - 16 : 0];
            else
                key = S0[i];
            const size_t o = orow * kout + i;
            const uint32_t id = (uint32_t)key;
            if (tp.idx_is_int64)
                reinterpret_cast<long long *>(tp.idx)[o] = (long long)id;
            else
                reinterpret_cast<int *>(tp.idx)[o] = (int)id;
            if (tp.dist) tp.dist[o] = sortable2f((uint32_t)(key >> 32));
        }
    }
}

// Queries to redo exactly, known as soon as the scan has published the list lengths: fewer than k
// candidates below an estimated bound, or an overflowed list. Writes a flag per query, the query
// list and (once per warp = one 32-query tile) the tile list.
static __global__ void __launch_bounds__(TOPK_THREADS) knn_flag_kernel(int S, TopkParams tp) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z, tile = blockIdx.x;
    const int qi = tile * TOPK_THREADS + tid;
    const bool valid = qi < S;
    uint32_t total = 0u;
    bool overflow = false;
    for (int s = 0; s < tp.nsplit; ++s) {
        const size_t warp_linear = (size_t)(b * tp.nsplit + s) * tp.scan_tiles + tile;
        const uint32_t ntot = tp.cand_cnt[warp_linear * 128 + tid];
        overflow |= ntot > (uint32_t)tp.cap;
        total += min(ntot, (uint32_t)tp.cap);
    }
    const size_t qrow = (size_t)b * S + (valid ? qi : 0);
    const bool redo = valid && (total < (uint32_t)tp.kout || overflow);
    const int tiles_per_cloud = (S + 31) / 32;
    int *tile_list = tp.fail_list + (size_t)gridDim.z * S;
    int *query_list = tile_list + (size_t)gridDim.z * tiles_per_cloud;
    if (redo) {
        tp.fail_list[qrow] = 1;
        query_list[atomicAdd(tp.fail_count + 1, 1)] = (int)qrow;
    }
    if (__any_sync(0xffffffffu, redo) && lane == 0)
        tile_list[atomicAdd(tp.fail_count, 1)] = b * tiles_per_cloud + qi / 32;
}

}  // namespace b200pci
