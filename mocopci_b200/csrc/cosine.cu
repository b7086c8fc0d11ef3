// Feature-space cosine k-NN -- replaces models/pointconv_util.py:111-127
// (cosine_distance: normalise both clouds, dist = 1 - bmm) + :142-153 (knn_point_cosine: topk), which
// the model calls ~40 times per forward on 256..2048 points with 64..256 channels.
//
// This is the one deep contraction on the hot path, so it runs on the tensor cores:
//   cos_pack_kernel   x / sqrt(sum x^2 + 1e-8) per row (the reference's formula), each value split
//                     into two exact TF32 pieces (hi = low 13 mantissa bits cleared, lo = the
//                     next 11 bits), written as K-major no-swizzle operand blocks
//                     [row tile][16-channel chunk][hi|lo][4 x (128 rows x 4 floats)] = 16 KB each:
//                     one 1-D TMA bulk copy per block, in exactly the layout tcgen05.mma reads.
//   cos_knn_kernel    a CTA owns 128 queries x 256 refs (two 128-column FP32 accumulators in tensor
//                     memory) and loops over the channel chunks: warp 4 streams the operand blocks
//                     through a 3-stage shared-memory ring with TMA, warp 5 issues per chunk and
//                     ref tile 2 K-steps x 3 tcgen05.mma.kind::tf32 (hi*hi into the main accumulator,
//                     hi*lo + lo*hi into a separate CORRECTION accumulator: the tensor core truncates
//                     when it aligns addends, a bias of ~1/2 ulp of the running sum per MMA, so the
//                     2^-11-times smaller corrections are kept out of the main chain and added
//                     once, rounded to nearest, in the epilogue; the dropped lo*lo term is ~2^-22
//                     relative), and after
//                     the last chunk eight epilogue warps (TMEM lane quarter x ref tile: one query
//                     and 128 refs per thread) read the dot products with tcgen05.ld, form 1 - dot
//                     and keep the k smallest with the threshold / sorting-network fold of the
//                     Euclidean kernels (keys sortable(distance) << 32 | index: the lowest index wins
//                     ties); the two tiles' lists are merged through shared memory. The ref splits
//                     of a query tile (N / 256 CTAs) leave sorted partial lists; the LAST CTA of
//                     a tile to finish (atomic ticket) merges them with the same networks and writes
//                     the int64 indices.
// Two launches per call, no distance matrix, no separate normalise / rsub / topk passes.
//
// Parity: cuBLAS' FP32 summation order inside torch.bmm is unspecified, so this op cannot be
// bit-exact against the reference; distances agree to ~1e-6 absolute and the neighbour sets are
// equal wherever the reference's k-th and (k+1)-th distances are further apart than that
// (tests/test_gpu_parity.py::test_knn_point_cosine_*).
#include "nbr_scan_tc.cuh"

namespace b200pci {

constexpr int COS_ROWS = 128;                 // rows per operand tile (= MMA M and N)
constexpr int COS_CHUNK = 16;                 // channels per operand block (two K = 8 steps)
constexpr int COS_BLOCK_FLOATS = 2 * 4 * COS_ROWS * 4;  // [hi|lo][4 sub-chunks][128 rows][4]
constexpr uint32_t COS_BLOCK_BYTES = COS_BLOCK_FLOATS * sizeof(float);  // 16 KB
constexpr int COS_STAGES = 3;
constexpr int COS_REFS_PER_CTA = 2 * COS_ROWS;  // two accumulators
constexpr int COS_EPI_WARPS = 8;  // TMEM lane quarter (warp % 4) x ref tile (warp / 4)
constexpr int COS_THREADS = (COS_EPI_WARPS + 2) * 32;
constexpr uint32_t COS_TMEM_COLS = 512;  // per ref tile: main (hi*hi) and correction (hi*lo + lo*hi) accumulators
constexpr int COS_MAX_SPLIT = 16;

struct CosSmem {
    static constexpr size_t ring = (size_t)COS_STAGES * 3 * COS_BLOCK_BYTES;            // A + 2 B blocks
    static constexpr size_t buf = (size_t)COS_EPI_WARPS * 16 * 32 * sizeof(u64);        // candidate buffers
    // (the hand-over of the second ref tile's list to the first re-uses the operand ring: 32 KB)
    static constexpr size_t ctrl = 256;
    static constexpr size_t total = ring + buf + ctrl;
};

// rows of q (z < B) and of r (z >= B): norm, normalise, split, store. A CTA owns 32 rows; its 256
// threads = 32 rows x 8 channel groups (consecutive lanes = consecutive rows: coalesced for the
// channel-major [B,C,N] views the model passes), so that a 2048-point cloud is 64 CTAs per operand
// instead of 16. x(b, n, c) = base[b*sb + n*sn + c*sc].
constexpr int COS_PACK_ROWS = 32;
__global__ void __launch_bounds__(256)
    cos_pack_kernel(int B, int S, int N, int C, const float *__restrict__ q, long long q_sb, long long q_sn,
                    long long q_sc, const float *__restrict__ r, long long r_sb, long long r_sn, long long r_sc,
                    float *__restrict__ opq, float *__restrict__ opr, int qtiles, int rtiles,
                    unsigned int *__restrict__ tickets, int nticket) {
    __shared__ float part[8][COS_PACK_ROWS];
    const bool is_r = (int)blockIdx.z >= B;
    const int b = is_r ? blockIdx.z - B : blockIdx.z;
    const int tiles = is_r ? rtiles : qtiles;
    if (blockIdx.x == 0 && blockIdx.z == 0)
        for (int i = threadIdx.x; i < nticket; i += 256) tickets[i] = 0u;
    const int n0 = blockIdx.x * COS_PACK_ROWS;  // first row of this CTA
    if (n0 >= tiles * COS_ROWS) return;
    const int rows = is_r ? N : S;
    const float *src = is_r ? r : q;
    const long long sb = is_r ? r_sb : q_sb, sn = is_r ? r_sn : q_sn, sc = is_r ? r_sc : q_sc;
    const int nchunks = C / COS_CHUNK;
    const int row = threadIdx.x & 31, grp = threadIdx.x >> 5;  // 8 channel groups
    const int n = n0 + row;
    const bool valid = n < rows;
    const float *x = src + b * sb + (long long)n * sn;
    float s = 0.f;
    if (valid)
        for (int c = grp; c < C; c += 8) {
            const float v = x[c * sc];
            s = fmaf(v, v, s);
        }
    part[grp][row] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) tot = __fadd_rn(tot, part[g][row]);
    // models/pointconv_util.py:122-123: x / sqrt(sum(x ** 2) + 1e-8)
    const float nrm = __fsqrt_rn(__fadd_rn(tot, 1e-8f));
    const int tile = n / COS_ROWS, trow = n % COS_ROWS;
    float *op = (is_r ? opr : opq) + ((size_t)b * tiles + tile) * (size_t)nchunks * COS_BLOCK_FLOATS;
    // a thread writes 4-channel sub-chunks: (chunk, sub) pairs grp, grp + 8, ...
    for (int cs = grp; cs < nchunks * 4; cs += 8) {
        const int ch = cs >> 2, sub = cs & 3;
        float h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = valid ? __fdiv_rn(x[(ch * COS_CHUNK + sub * 4 + i) * sc], nrm) : 0.f;
            h[i] = tf32_hi(v);
            l[i] = tf32_hi(__fsub_rn(v, h[i]));
        }
        float4 *blk = reinterpret_cast<float4 *>(op + (size_t)ch * COS_BLOCK_FLOATS);
        blk[sub * COS_ROWS + trow] = make_float4(h[0], h[1], h[2], h[3]);
        blk[(4 + sub) * COS_ROWS + trow] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

struct CosArgs {
    int S, N, C, k, nsplit, qtiles, rtiles;
    const float *opq, *opr;
    unsigned long long *part;  // [B*S][nsplit][K] sorted partial lists (nsplit > 1)
    unsigned int *tickets;     // [B*qtiles]
    void *idx;                 // [B,S,k]
    int idx_is_int64;
    float *dist;               // nullable [B,S,k]
};

template <int K>
__global__ void __launch_bounds__(COS_THREADS, 1) cos_knn_kernel(CosArgs a) {
    constexpr int NBLK = K / 16;
    static_assert(K == 16 || K == 32, "cos_knn_kernel: K = 16 or 32");
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *ring = smem;
    u64 *cbuf = reinterpret_cast<u64 *>(smem + CosSmem::ring);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + CosSmem::ring + CosSmem::buf);
    uint64_t *full = bars, *empty = bars + COS_STAGES, *acc_full = bars + 2 * COS_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * COS_STAGES + 1);
    int *s_last = reinterpret_cast<int *>(tmem_slot + 1);
    const int qt = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
    const int nchunks = a.C / COS_CHUNK;

    if (tid == 0) {
        for (int s = 0; s < COS_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == COS_EPI_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(COS_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == COS_EPI_WARPS) {
        // ---- TMA producer: per chunk the query block and the two ref blocks ----
        if (lane == 0) {
            const float *qa = a.opq + ((size_t)b * a.qtiles + qt) * (size_t)nchunks * COS_BLOCK_FLOATS;
            const float *r0 = a.opr + ((size_t)b * a.rtiles + 2 * split) * (size_t)nchunks * COS_BLOCK_FLOATS;
            const float *r1 = r0 + (size_t)nchunks * COS_BLOCK_FLOATS;
            for (int ch = 0; ch < nchunks; ++ch) {
                const int s = ch % COS_STAGES;
                if (ch >= COS_STAGES) mbar_wait_suspend(&empty[s], ((ch / COS_STAGES) - 1) & 1);
                unsigned char *st = ring + (size_t)s * 3 * COS_BLOCK_BYTES;
                mbar_arrive_expect_tx(&full[s], 3 * COS_BLOCK_BYTES);
                tma_load_1d(st, qa + (size_t)ch * COS_BLOCK_FLOATS, COS_BLOCK_BYTES, &full[s]);
                tma_load_1d(st + COS_BLOCK_BYTES, r0 + (size_t)ch * COS_BLOCK_FLOATS, COS_BLOCK_BYTES, &full[s]);
                tma_load_1d(st + 2 * COS_BLOCK_BYTES, r1 + (size_t)ch * COS_BLOCK_FLOATS, COS_BLOCK_BYTES, &full[s]);
            }
        }
    } else if (warp == COS_EPI_WARPS + 1) {
        // ---- MMA issuer: dot += qh.rh + qh.rl + ql.rh over two K = 8 steps per chunk ----
        if (lane == 0) {
            constexpr uint32_t KSTEP = 2 * TC_KCHUNK_BYTES;  // two 4-float sub-chunks
            constexpr uint32_t LO = 4 * TC_KCHUNK_BYTES;     // the lo half of a block
            for (int ch = 0; ch < nchunks; ++ch) {
                const int s = ch % COS_STAGES;
                mbar_wait_suspend(&full[s], (ch / COS_STAGES) & 1);
                tc_fence_after();
                const uint32_t sa = smem_u32(ring + (size_t)s * 3 * COS_BLOCK_BYTES);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t sbj = sa + (1 + j) * COS_BLOCK_BYTES;
                    const uint32_t d = tmem_base + j * COS_ROWS, dc = d + 2 * COS_ROWS;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t ah = sa + ks * KSTEP, al = ah + LO, bh = sbj + ks * KSTEP, bl = bh + LO;
                        tc_mma(d, tc_smem_desc(ah), tc_smem_desc(bh), (ch | ks) ? 1u : 0u);
                        tc_mma(dc, tc_smem_desc(ah), tc_smem_desc(bl), (ch | ks) ? 1u : 0u);
                        tc_mma(dc, tc_smem_desc(al), tc_smem_desc(bh), 1u);
                    }
                }
                tc_commit(&empty[s]);
            }
            tc_commit(acc_full);
        }
    } else {
        // ---- epilogue: one query (TMEM lane) and one ref tile (128 columns) per thread ----
        const int quarter = warp & 3, tile = warp >> 2;
        const int row = quarter * 32 + lane;
        const int qi = qt * COS_ROWS + row;
        const bool live = qi < a.S;
        u64 *buf = cbuf + (size_t)warp * 16 * 32 + lane;  // [16][32]
        u64 S0[16], S1[NBLK > 1 ? 16 : 1];
#pragma unroll
        for (int i = 0; i < 16; ++i) S0[i] = B200PCI_KEY_INF;
        if constexpr (NBLK > 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) S1[i] = B200PCI_KEY_INF;
        }
        auto merge_sorted16 = [&](u64 (&Cn)[16]) {
            if constexpr (NBLK == 1) {
                merge_low16(S0, Cn);
            } else {
                merge_low16(S1, Cn);
                merge_full16(S0, S1);
            }
        };
        const int kl = a.k - 1;
        float tcur = live ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
        int nb = 0;
        auto fold = [&]() {
            u64 Cn[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) Cn[i] = (i < nb) ? buf[i * 32] : ~0ull;
            nb = 0;
            sort16(Cn);
            merge_sorted16(Cn);
            u64 kth;
            if constexpr (NBLK == 1)
                kth = sel16(S0, kl);
            else
                kth = (kl < 16) ? sel16(S0, kl) : sel16(S1, kl - 16);
            tcur = fminf(tcur, sortable2f((uint32_t)(kth >> 32)));
        };
        mbar_wait_suspend(acc_full, 0);
        __syncwarp();
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + tile * COS_ROWS;
        const int n_base = split * COS_REFS_PER_CTA + tile * COS_ROWS;
#pragma unroll 1
        for (int c = 0; c < COS_ROWS / 32; ++c) {
            if (n_base + c * 32 >= a.N) break;  // (uniform) nothing but padding from here on
            float v[32], vc[32];
            tc_ld32(trow + c * 32, v);
            tc_ld32(trow + 2 * COS_ROWS + c * 32, vc);
            tc_ld_wait(v);
            tc_ld_pin(vc);
#pragma unroll
            for (int i0 = 0; i0 < 32; i0 += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = n_base + c * 32 + i0 + u;
                    const float d = __fsub_rn(1.0f, __fadd_rn(v[i0 + u], vc[i0 + u]));  // pointconv_util.py:125
                    if (n < a.N && d < tcur) {
                        buf[nb * 32] = make_key(d, (uint32_t)n);
                        ++nb;
                    }
                }
                if (__any_sync(0xffffffffu, nb > 12)) fold();
            }
        }
        if (__any_sync(0xffffffffu, nb > 0)) fold();
        tc_fence_before();
        // tile 1 hands its list to tile 0 through the (now idle) operand ring
        u64 *xch = reinterpret_cast<u64 *>(ring) + row;  // [K][128]
        if (tile == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) xch[i * COS_ROWS] = S0[i];
            if constexpr (NBLK > 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) xch[(16 + i) * COS_ROWS] = S1[i];
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(COS_EPI_WARPS * 32) : "memory");
        if (tile == 0) {
#pragma unroll 1
            for (int blk = 0; blk < NBLK; ++blk) {
                u64 Cn[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) Cn[i] = xch[(blk * 16 + i) * COS_ROWS];
                merge_sorted16(Cn);
            }
        }
        if (live && tile == 0) {
            const size_t qrow = (size_t)b * a.S + qi;
            if (a.nsplit == 1) {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (i < a.k) {
                        u64 key;
                        if constexpr (NBLK > 1)
                            key = (i < 16) ? S0[i < 16 ? i : 0] : S1[i >= 16 ? i - 16 : 0];
                        else
                            key = S0[i];
                        const size_t o = qrow * a.k + i;
                        if (a.idx_is_int64)
                            reinterpret_cast<long long *>(a.idx)[o] = (long long)(uint32_t)key;
                        else
                            reinterpret_cast<int *>(a.idx)[o] = (int)(uint32_t)key;
                        if (a.dist) a.dist[o] = sortable2f((uint32_t)(key >> 32));
                    }
                }
            } else {
                // partial lists are stored K wide ([split][K], unused slots = +inf keys) so that the
                // merge reads whole sorted 16-blocks
                u64 *dst = a.part + (qrow * a.nsplit + split) * K;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if constexpr (NBLK > 1)
                        dst[i] = (i < 16) ? S0[i < 16 ? i : 0] : S1[i >= 16 ? i - 16 : 0];
                    else
                        dst[i] = S0[i];
                }
            }
        }
    }
    // ---- the last CTA of this query tile merges the nsplit partial lists ----
    if (a.nsplit > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int t = atomicAdd(&a.tickets[b * a.qtiles + qt], 1u);
            *s_last = (t == (unsigned int)a.nsplit - 1u);
        }
        __syncthreads();
        if (*s_last && warp < COS_EPI_WARPS) {
            // two threads per query row: each folds one half of the splits' partial lists, the
            // second hands its result over through the (idle) operand ring, the first folds it in
            // and writes the row (one thread walking all nsplit lists took ~17 us of the call)
            __threadfence();
            const int h = warp >> 2;
            const int row = (warp & 3) * 32 + lane;
            const int qi = qt * COS_ROWS + row;
            const bool live = qi < a.S;
            const size_t qrow = (size_t)b * a.S + (live ? qi : 0);
            const int s_begin = h ? a.nsplit / 2 : 0, s_end = h ? a.nsplit : a.nsplit / 2;
            const u64 *src = a.part + qrow * a.nsplit * K;
            u64 S0[16], S1[NBLK > 1 ? 16 : 1];
            auto fold = [&](u64(&Cn)[16]) {
                if constexpr (NBLK == 1) {
                    merge_low16(S0, Cn);
                } else {
                    merge_low16(S1, Cn);
                    merge_full16(S0, S1);
                }
            };
#pragma unroll
            for (int i = 0; i < 16; ++i) S0[i] = live ? __ldcg(src + (size_t)s_begin * K + i) : ~0ull;
            if constexpr (NBLK > 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) S1[i] = live ? __ldcg(src + (size_t)s_begin * K + 16 + i) : ~0ull;
            }
#pragma unroll 1
            for (int sp = s_begin + 1; sp < s_end; ++sp) {
#pragma unroll 1
                for (int blk = 0; blk < NBLK; ++blk) {
                    u64 Cn[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) Cn[i] = live ? __ldcg(src + (size_t)sp * K + blk * 16 + i) : ~0ull;
                    fold(Cn);
                }
            }
            u64 *xch = reinterpret_cast<u64 *>(ring) + row;  // [K][128]
            if (h == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) xch[i * COS_ROWS] = S0[i];
                if constexpr (NBLK > 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) xch[(16 + i) * COS_ROWS] = S1[i];
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(COS_EPI_WARPS * 32) : "memory");
            if (h == 0) {
#pragma unroll 1
                for (int blk = 0; blk < NBLK; ++blk) {
                    u64 Cn[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) Cn[i] = xch[(blk * 16 + i) * COS_ROWS];
                    fold(Cn);
                }
                if (live) {
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        if (i < a.k) {
                            u64 key;
                            if constexpr (NBLK > 1)
                                key = (i < 16) ? S0[i < 16 ? i : 0] : S1[i >= 16 ? i - 16 : 0];
                            else
                                key = S0[i];
                            const size_t o = qrow * a.k + i;
                            if (a.idx_is_int64)
                                reinterpret_cast<long long *>(a.idx)[o] = (long long)(uint32_t)key;
                            else
                                reinterpret_cast<int *>(a.idx)[o] = (int)(uint32_t)key;
                            if (a.dist) a.dist[o] = sortable2f((uint32_t)(key >> 32));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == COS_EPI_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COS_TMEM_COLS) : "memory");
    }
}

struct CosPlan {
    int qtiles, rtiles, nsplit, nchunks;
    size_t opq_bytes, opr_bytes, part_bytes, ticket_bytes;
    size_t total() const { return opq_bytes + opr_bytes + part_bytes + ticket_bytes; }
};

static bool cos_plan(int B, int S, int N, int C, int k, CosPlan &pl) {
    if (B <= 0 || S <= 0 || N <= 0 || C < COS_CHUNK || C % COS_CHUNK != 0 || C > 1024 || k < 1 || k > 32 || k > N)
        return false;
    pl.nsplit = ceil_div(N, COS_REFS_PER_CTA);
    if (pl.nsplit > COS_MAX_SPLIT || B > 65535) return false;
    pl.qtiles = ceil_div(S, COS_ROWS);
    pl.rtiles = 2 * pl.nsplit;
    pl.nchunks = C / COS_CHUNK;
    pl.opq_bytes = align_up((size_t)B * pl.qtiles * pl.nchunks * COS_BLOCK_BYTES, 256);
    pl.opr_bytes = align_up((size_t)B * pl.rtiles * pl.nchunks * COS_BLOCK_BYTES, 256);
    pl.part_bytes = pl.nsplit > 1 ? align_up((size_t)B * S * pl.nsplit * (k <= 16 ? 16 : 32) * sizeof(u64), 256) : 0;
    pl.ticket_bytes = align_up((size_t)B * pl.qtiles * sizeof(unsigned int), 256);
    return true;
}

}  // namespace b200pci

using namespace b200pci;

extern "C" size_t b200pci_knn_cosine_workspace_bytes(int B, int S, int N, int C, int k) {
    CosPlan pl;
    return cos_plan(B, S, N, C, k, pl) ? pl.total() : 0;
}

extern "C" int b200pci_knn_cosine(int B, int S, int N, int C, int k, const float *q, int64_t q_sb, int64_t q_sn,
                                  int64_t q_sc, const float *r, int64_t r_sb, int64_t r_sn, int64_t r_sc,
                                  void *idx, int idx_is_int64, float *dist, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200PCI_CHECK_ARG(B >= 0 && S >= 0 && N >= 0, "knn_cosine: negative size");
    if (B == 0 || S == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(k >= 1 && k <= N, "selected index k out of range (k=%d, N=%d)", k, N);
    CosPlan pl;
    B200PCI_CHECK_ARG(cos_plan(B, S, N, C, k, pl),
                      "knn_cosine: unsupported shape (needs C %% 16 == 0, C <= 1024, k <= 32, N <= %d)",
                      COS_MAX_SPLIT * COS_REFS_PER_CTA);
    B200PCI_CHECK_ARG(q && r && idx, "knn_cosine: null pointer");
    if (!workspace || workspace_bytes < pl.total() || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
        set_error("knn_cosine: workspace of %zu bytes (256-B aligned) required, got %zu", pl.total(), workspace_bytes);
        return B200PCI_EWORKSPACE;
    }
    char *ws = reinterpret_cast<char *>(workspace);
    float *opq = reinterpret_cast<float *>(ws);
    float *opr = reinterpret_cast<float *>(ws + pl.opq_bytes);
    unsigned long long *part = reinterpret_cast<unsigned long long *>(ws + pl.opq_bytes + pl.opr_bytes);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(ws + pl.opq_bytes + pl.opr_bytes + pl.part_bytes);
    const int maxt = (pl.qtiles > pl.rtiles ? pl.qtiles : pl.rtiles) * (COS_ROWS / COS_PACK_ROWS);
    cos_pack_kernel<<<dim3(maxt, 1, 2 * B), 256, 0, st>>>(B, S, N, C, q, q_sb, q_sn, q_sc, r, r_sb, r_sn, r_sc, opq,
                                                         opr, pl.qtiles, pl.rtiles, tickets, B * pl.qtiles);
    B200PCI_LAUNCH_CHECK("cos_pack_kernel");
    CosArgs a;
    a.S = S;
    a.N = N;
    a.C = C;
    a.k = k;
    a.nsplit = pl.nsplit;
    a.qtiles = pl.qtiles;
    a.rtiles = pl.rtiles;
    a.opq = opq;
    a.opr = opr;
    a.part = part;
    a.tickets = tickets;
    a.idx = idx;
    a.idx_is_int64 = idx_is_int64;
    a.dist = dist;
    dim3 grid(pl.qtiles, pl.nsplit, B);
    if (k <= 16) {
        auto kern = cos_knn_kernel<16>;
        B200PCI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CosSmem::total));
        kern<<<grid, COS_THREADS, CosSmem::total, st>>>(a);
    } else {
        auto kern = cos_knn_kernel<32>;
        B200PCI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CosSmem::total));
        kern<<<grid, COS_THREADS, CosSmem::total, st>>>(a);
    }
    B200PCI_LAUNCH_CHECK("cos_knn_kernel");
    return B200PCI_OK;
}
