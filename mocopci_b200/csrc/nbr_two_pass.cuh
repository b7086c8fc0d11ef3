// Two-pass top-k for big clouds with a known admission bound (DESIGN.md "Neighbourhood engine"):
//
//   scan   (nbr_scan)          the streaming filter of nbr_engine.cuh with nothing else in the
//                              kernel: a warp owns 4 x 32 queries and one split of the refs, and
//                              appends one (step << 8 | 8-group mask) entry per flagged 32-ref step
//                              to the query's pending list in global memory (predicated store, no
//                              branch). No drains, no selection state: the loop is the kernel.
//   select (knn_select_kernel) one THREAD per (query, half of the splits) walks its lists,
//                              re-evaluates the flagged groups in the exact reference arithmetic
//                              (refs gathered from the 64-byte group records in L2), buffers the
//                              candidates below the bound and folds them 16 at a time into a sorted
//                              best-K held in registers (sorting networks). 128-thread CTAs: the L2
//                              gathers of one warp hide behind the arithmetic of the others.
//
// The bound comes from the threshold pre-pass (knn_tau_kernel): an estimate for k >= 5 (queries
// that end with fewer than k candidates, or whose list overflowed, go to the exact redo kernel), a
// guaranteed bound for k <= 4.
#pragma once
#include "nbr_engine.cuh"

namespace b200pci {

constexpr int SCAN_CAP = 96;  // pending entries per (query, split); a multiple of 4

// ---- pass 1: scan ----------------------------------------------------------------------------
template <int CW, int STAGES>
__device__ __forceinline__ void nbr_scan(const NbrParams &p) {
    constexpr int QT = NBR_QT;
    constexpr int G4 = NBR_TILE / 4;  // float4 per row per stage
    constexpr size_t warp_ring_bytes = (size_t)STAGES * 4 * NBR_TILE * sizeof(float);
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *tiles = reinterpret_cast<float *>(smem + (size_t)warp * warp_ring_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + warp_ring_bytes * CW) + warp * STAGES;

    const int b = blockIdx.z, split = blockIdx.y;
    const size_t warp_linear =
        ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * CW + warp;
    const int tile0 = split * p.tiles_per_split;
    const int ntiles = min(p.tiles_per_split, p.total_tiles - tile0);
    const float *ws = p.ws_ref + (size_t)b * 4 * p.Npad;
    constexpr uint32_t stage_bytes = 4 * NBR_TILE * sizeof(float);

    auto issue_tile = [&](int t) {  // lane 0 of the owning warp
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
    __syncwarp();

    QueryRegs q[QT];
    float thr[QT];
    int cnt[QT];
    const int qi0 = (blockIdx.x * CW + warp) * (QT * 32) + lane;  // slot j: + 32 * j
    uint32_t *pend = p.pend + warp_linear * (size_t)(QT * SCAN_CAP * 32) + lane * 4;
#pragma unroll
    for (int j = 0; j < QT; ++j) {
        const int qi = qi0 + 32 * j;
        float x = 0.f, y = 0.f, z = 0.f, t0 = __int_as_float(0xff800000);  // -inf: never flagged
        if (qi < p.S) {
            const float *src = p.q + b * p.q_sb + qi * p.q_sp;
            x = src[0];
            y = src[p.q_sc];
            z = src[2 * p.q_sc];
            t0 = p.tau_in ? p.tau_in[(size_t)b * p.S + qi] : p.tau_uniform;
        }
        q[j].set(x, y, z);
        thr[j] = q[j].threshold(t0);
        cnt[j] = 0;
    }

    constexpr int SPT = G4 / NBR_BLK;  // steps per tile
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (t / STAGES) & 1);
        const float4 *sX = reinterpret_cast<const float4 *>(tiles + (size_t)(s * 4) * NBR_TILE);
        float4 X = sX[0], Y = sX[G4], Z = sX[2 * G4], W = sX[3 * G4];
        uint32_t ent = (uint32_t)((tile0 + t) * SPT) << 8;
#pragma unroll 1
        for (int g0 = 0; g0 < G4; g0 += NBR_BLK) {
            const float4 *gX = sX + g0;
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
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                if (m8[j] != 0u && cnt[j] < SCAN_CAP)
                    pend[j * (SCAN_CAP * 32) + nbr_pend_off(cnt[j])] = ent | m8[j];
                if (m8[j] != 0u) ++cnt[j];  // keeps counting past the capacity: overflow marker
            }
            ent += 1u << 8;
        }
        // this warp is done with the stage: refill it with the tile STAGES ahead
        __syncwarp();
        if (lane == 0 && t + STAGES < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile(t + STAGES);
        }
    }
    uint32_t *pc = p.pend_cnt + warp_linear * (QT * 32) + lane;
#pragma unroll
    for (int j = 0; j < QT; ++j) pc[j * 32] = (uint32_t)cnt[j];
}

// ---- pass 2: select --------------------------------------------------------------------------
struct SelectParams {
    void *idx;    // int64/int32 [B,S,kout]
    float *dist;  // nullable
    int idx_is_int64;
    int kout;
    int *fail_count;  // [0] flagged tiles, [1] flagged queries (pre-zeroed)
    int *fail_list;   // [B*S] redo flags (pre-zeroed: under-filled, or a pending list overflowed),
                      // then the list of 32-query tiles with a flagged query [B*ceil(S/32)],
                      // then the list of flagged queries [B*S]
    int scan_tiles;  // query tiles of the scan grid (gridDim.x of the scan)
};

constexpr int SEL_THREADS = 128;
constexpr int SEL_BUF = 32;  // candidate buffer depth per thread (shared memory)

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// A CTA owns SEL_Q = 64 consecutive queries of one cloud; thread (h, ql) = (tid / 64, tid % 64)
// walks the lists of query ql for the splits h, h + 2, h + 4, ... (two half-length dependent chains
// per query instead of one), then the h = 1 threads hand their sorted best-K to their h = 0
// partners through shared memory, which merge and write the result.
constexpr int SEL_Q = SEL_THREADS / 2;

template <int MODE, int K>
__global__ void __launch_bounds__(SEL_THREADS)
    knn_select_kernel(NbrParams p, SelectParams sp) {
    constexpr bool NET = K > 4;
    constexpr int NBLK = NET ? K / 16 : 1;
    constexpr int KR = NET ? 16 : K;
    static_assert(NBLK <= 2, "select kernel: K <= 32");
    // candidate buffers [SEL_BUF][SEL_THREADS]; reused at the end for the hand-over [K][SEL_Q]
    __shared__ u64 buf_s[NET ? SEL_BUF * SEL_THREADS : K * SEL_Q];
    // the next two quads of every thread's list, fetched with cp.async: no register is the
    // destination of a load another lane issued, so lanes at different list positions never
    // wait for each other's entry loads
    __shared__ uint4 quad_s[2][SEL_THREADS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int h = tid / SEL_Q, ql = tid % SEL_Q;
    const int b = blockIdx.z;
    const int qi = blockIdx.x * SEL_Q + ql;
    const bool valid = qi < p.S;
    const int tile = qi / (NBR_QT * 32), j = (qi / 32) % NBR_QT;  // scan warp / slot of this query
    u64 *buf = buf_s + tid;

    QueryRegs q;
    float tau = __int_as_float(0xff800000);
    {
        float x = 0.f, y = 0.f, z = 0.f;
        if (valid) {
            const float *src = p.q + b * p.q_sb + qi * p.q_sp;
            x = src[0];
            y = src[p.q_sc];
            z = src[2 * p.q_sc];
            tau = p.tau_in[(size_t)b * p.S + qi];
        }
        q.set(x, y, z);
    }
    const float *grp = p.ws_grp + (size_t)b * 4 * p.Npad;

    // sorted best-K in registers: blocks of 16 (block 0 = smallest), or K <= 4 keys
    u64 S0[KR], S1[NBLK > 1 ? 16 : 1];
#pragma unroll
    for (int i = 0; i < KR; ++i) S0[i] = B200PCI_KEY_INF;
    if constexpr (NBLK > 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) S1[i] = B200PCI_KEY_INF;
    }
    int nb = 0;
    float tcur = tau;  // admission: d < tcur (tightened by the folds)
    const int kl = sp.kout - 1;

    auto merge_sorted16 = [&](u64 (&C)[16]) {  // fold an ascending chunk of 16 into the best-K
        if constexpr (NET) {
            if constexpr (NBLK == 1) {
                merge_low16(S0, C);
            } else {
                merge_low16(S1, C);    // S1 = 16 smallest of (top block U chunk)
                merge_full16(S0, S1);  // S0 = low half, S1 = high half
            }
        }
    };
    auto fold16 = [&](int first) {
        if constexpr (NET) {
            u64 C[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
                C[i] = (first + i < nb) ? buf[(first + i) * SEL_THREADS] : ~0ull;
            sort16(C);
            merge_sorted16(C);
            u64 kth;
            if constexpr (NBLK == 1)
                kth = sel16(S0, kl);
            else
                kth = (kl < 16) ? sel16(S0, kl) : sel16(S1, kl - 16);
            tcur = fminf(tcur, sortable2f((uint32_t)(kth >> 32)));
        }
    };
    auto fold_all = [&]() {
        fold16(0);
        if (__any_sync(0xffffffffu, nb > 16)) fold16(16);
        nb = 0;
    };

    bool overflow = false;
    for (int s = h; s < p.nsplit; s += 2) {
        const size_t warp_linear = ((size_t)(b * p.nsplit + s) * sp.scan_tiles + tile);
        int cnt = valid ? (int)p.pend_cnt[warp_linear * (NBR_QT * 32) + j * 32 + lane] : 0;
        if (cnt > SCAN_CAP) {
            overflow = true;
            cnt = SCAN_CAP;
        }
        const uint4 *quads = reinterpret_cast<const uint4 *>(
            p.pend + warp_linear * (size_t)(NBR_QT * SCAN_CAP * 32) + j * (SCAN_CAP * 32) + lane * 4);
        auto prefetch = [&](int qd) {
            if (qd * 4 < cnt) cp_async16(&quad_s[qd & 1][tid], quads + (size_t)qd * 32);
            cp_async_commit();
        };
        prefetch(0);
        prefetch(1);
        int e = 0;
        uint32_t m8 = 0u, gs = 0u;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        // every lane walks its own list (entry e, remaining mask m8): advance() yields the lane's
        // next flagged group. A round evaluates one group per lane that still has one; its
        // 64-byte record comes in two 256-bit loads (every lane gathers a different record: the
        // cost is the number of requests, and the other warps of the SM hide the latency).
        auto advance = [&](bool &has) -> uint32_t {
            if (m8 == 0u && e < cnt) {
                const int k = e & 3;
                if (k == 0) {
                    cp_async_wait<1>();  // this thread's quad e/4 has landed
                    w = quad_s[(e >> 2) & 1][tid];
                    prefetch((e >> 2) + 2);
                }
                const uint32_t ent = (k == 0) ? w.x : (k == 1) ? w.y : (k == 2) ? w.z : w.w;
                m8 = ent & 0xffu;
                gs = ent >> 8;
                ++e;
            }
            has = m8 != 0u;
            const int bit = has ? (31 - __clz((int)m8)) : 0;  // highest bit = lowest group
            m8 &= ~(1u << bit);
            return has ? gs * NBR_BLK + (uint32_t)(7 - bit) : 0u;
        };
        while (true) {
            bool has;
            const uint32_t gid = advance(has);
            if (!__any_sync(0xffffffffu, has)) break;
            float4 X, Y, Z, Wc;
            ldg256(grp + (size_t)gid * 16, X, Y);
            ldg256(grp + (size_t)gid * 16 + 8, Z, Wc);
            float d[4];
            dist4n<MODE>(q, X, Y, Z, Wc, gid * 4u, p.N, d);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool hit = has && d[i] < tcur;
                if constexpr (NET) {
                    if (hit) {
                        buf[nb * SEL_THREADS] = make_key(d[i], gid * 4u + i);
                        ++nb;
                    }
                } else {
                    if (__any_sync(0xffffffffu, hit)) {
                        u64 key = hit ? make_key(d[i], gid * 4u + i) : ~0ull;
#pragma unroll
                        for (int r = 0; r < K; ++r) ce64(S0[r], key);
                        u64 kth = S0[0];
#pragma unroll
                        for (int r = 1; r < K; ++r) kth = (kl == r) ? S0[r] : kth;
                        tcur = fminf(tcur, sortable2f((uint32_t)(kth >> 32)));
                    }
                }
            }
            if constexpr (NET) {
                if (__any_sync(0xffffffffu, nb > SEL_BUF - 4)) fold_all();
            }
        }
        cp_async_wait<0>();  // nothing of this split may land after the next one starts
    }
    if constexpr (NET) {
        if (__any_sync(0xffffffffu, nb > 0)) fold_all();
    }

    // hand-over: h = 1 -> h = 0 through shared memory (the candidate buffers are free now)
    __shared__ int over_s[SEL_Q];
    __syncthreads();
    u64 *xch = buf_s + ql;  // [K][SEL_Q]
    if (h == 1) {
#pragma unroll
        for (int i = 0; i < KR; ++i) xch[i * SEL_Q] = S0[i];
        if constexpr (NBLK > 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) xch[(16 + i) * SEL_Q] = S1[i];
        }
        over_s[ql] = overflow;
    }
    __syncthreads();
    if (h == 1) return;  // (whole warps)
    overflow |= over_s[ql] != 0;
    if constexpr (NET) {
#pragma unroll 1
        for (int blk = 0; blk < NBLK; ++blk) {
            u64 C[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) C[i] = xch[(blk * 16 + i) * SEL_Q];
            merge_sorted16(C);
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            u64 key = xch[i * SEL_Q];
#pragma unroll
            for (int r = 0; r < K; ++r) ce64(S0[r], key);
        }
    }

    const size_t qrow = (size_t)b * p.S + (valid ? qi : 0);
    const int kout = sp.kout;
    bool under = false;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        if (i < kout && valid) {
            u64 key;
            if constexpr (NBLK > 1)
                key = (i < 16) ? S0[i < 16 ? i : 0] : S1[i >= 16 ? i - 16 : 0];
            else
                key = S0[i];
            const size_t o = qrow * kout + i;
            const uint32_t id = (uint32_t)key;
            if (sp.idx_is_int64)
                reinterpret_cast<long long *>(sp.idx)[o] = (long long)id;
            else
                reinterpret_cast<int *>(sp.idx)[o] = (int)id;
            if (sp.dist) sp.dist[o] = sortable2f((uint32_t)(key >> 32));
            if (i == kout - 1) under = key >= B200PCI_KEY_INF;
        }
    }
    // queries to redo exactly: a flag per query, and (once per warp = one 32-query tile) the tile
    const bool redo = valid && (under || overflow);
    const int tiles_per_cloud = (p.S + 31) / 32;
    int *tile_list = sp.fail_list + (size_t)gridDim.z * p.S;
    int *query_list = tile_list + (size_t)gridDim.z * tiles_per_cloud;
    if (redo) {
        sp.fail_list[qrow] = 1;
        query_list[atomicAdd(sp.fail_count + 1, 1)] = (int)qrow;
    }
    if (__any_sync(0xffffffffu, redo) && lane == 0)
        tile_list[atomicAdd(sp.fail_count, 1)] = b * tiles_per_cloud + qi / 32;
}


// ---- ball query on the two-pass path -----------------------------------------------------------
// pointnet2/src/ball_query_gpu.cu:30-44: the first `nsample` refs (ascending index) with
// d2 < radius^2; the first hit fills every slot first; idx is pre-zeroed by the caller.
// The scan runs with the uniform bound radius^2; one thread per query then walks its lists in
// order and stops at nsample hits. A query whose list overflowed before nsample hits were
// found (possible only with > SCAN_CAP flagged steps in one split) is redone by a warp that scans
// the whole cloud in order.
struct BallSelectParams {
    int *idx;  // [B,S,nsample]
    int nsample;
    float radius2;
    int *fail_count;
    int *fail_list;
    int scan_tiles;
    int force_redo;  // test hook: send every query through the exact redo kernel
};

__global__ void __launch_bounds__(128) ball_select_kernel(NbrParams p, BallSelectParams sp) {
    const int tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
    const int b = blockIdx.z;
    const int qi = blockIdx.x * 128 + tid;
    const bool valid = qi < p.S;
    QueryRegs q;
    {
        float x = 0.f, y = 0.f, z = 0.f;
        if (valid) {
            const float *src = p.q + b * p.q_sb + qi * p.q_sp;
            x = src[0];
            y = src[p.q_sc];
            z = src[2 * p.q_sc];
        }
        q.set(x, y, z);
    }
    const float *grp = p.ws_grp + (size_t)b * 4 * p.Npad;
    const int ns = sp.nsample;
    const float r2 = sp.radius2;
    int *row = sp.idx + ((size_t)b * p.S + (valid ? qi : 0)) * ns;
    int found = valid ? 0 : ns;
    bool overflow = false;
    for (int s = 0; s < p.nsplit; ++s) {
        const size_t warp_linear = ((size_t)(b * p.nsplit + s) * sp.scan_tiles + blockIdx.x);
        int cnt = valid ? (int)p.pend_cnt[warp_linear * (NBR_QT * 32) + j * 32 + lane] : 0;
        if (cnt > SCAN_CAP) {
            overflow |= found < ns;
            cnt = SCAN_CAP;
        }
        const uint32_t *list =
            p.pend + warp_linear * (size_t)(NBR_QT * SCAN_CAP * 32) + j * (SCAN_CAP * 32) + lane * 4;
        int e = 0;
        uint32_t m8 = 0u, gs = 0u;
        while (true) {
            if (m8 == 0u && e < cnt && found < ns) {
                const uint32_t ent = list[nbr_pend_off(e)];
                m8 = ent & 0xffu;
                gs = ent >> 8;
                ++e;
            }
            const bool has = m8 != 0u && found < ns;
            if (!__any_sync(0xffffffffu, has)) break;
            const int bit = has ? (31 - __clz((int)m8)) : 0;  // highest bit = lowest group
            m8 &= ~(1u << bit);
            const uint32_t gid = has ? gs * NBR_BLK + (uint32_t)(7 - bit) : 0u;
            float4 X, Y, Z, Wn;
            ldg256(grp + (size_t)gid * 16, X, Y);
            ldg256(grp + (size_t)gid * 16 + 8, Z, Wn);
            float d[4];
            dist4<B200PCI_DIST_DIRECT>(q, X, Y, Z, gid * 4u, p.N, d);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (has && d[i] < r2 && found < ns) {
                    if (found == 0)
                        for (int l = 0; l < ns; ++l) row[l] = (int)(gid * 4u + i);
                    row[found] = (int)(gid * 4u + i);
                    ++found;
                }
            }
        }
    }
    if (valid && ((overflow && found < ns) || sp.force_redo))
        sp.fail_list[atomicAdd(sp.fail_count, 1)] = (int)((size_t)b * p.S + qi);
}

// exact redo: one warp per failed query scans the packed rows in index order
__global__ void __launch_bounds__(128) ball_fallback_kernel(NbrParams p, BallSelectParams sp) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int nfail = *sp.fail_count;
    const int ns = sp.nsample;
    for (int f = wid; f < nfail; f += nw) {
        const int qrow = sp.fail_list[f];
        const int b = qrow / p.S, qi = qrow - b * p.S;
        const float *src = p.q + b * p.q_sb + qi * p.q_sp;
        const float qx = src[0], qy = src[p.q_sc], qz = src[2 * p.q_sc];
        const float *ws = p.ws_ref + (size_t)b * 4 * p.Npad;
        int *row = sp.idx + (size_t)qrow * ns;
        int found = 0;
        for (int base = 0; base < p.N && found < ns; base += 32) {
            const int k = base + lane;
            bool hit = false;
            if (k < p.N) {
                const float dx = __fsub_rn(ws[k], qx), dy = __fsub_rn(ws[p.Npad + k], qy),
                            dz = __fsub_rn(ws[2 * (size_t)p.Npad + k], qz);
                hit = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))) < sp.radius2;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m == 0u) continue;
            if (found == 0) {
                const int first = base + __ffs(m) - 1;
                for (int l = lane; l < ns; l += 32) row[l] = first;
                __syncwarp();
            }
            const int pos = found + __popc(m & ((1u << lane) - 1u));
            if (hit && pos < ns) row[pos] = k;
            found += __popc(m);
        }
        __syncwarp();
    }
}

}  // namespace b200pci
