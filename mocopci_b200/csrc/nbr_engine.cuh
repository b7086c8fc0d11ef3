// The neighbourhood engine: one streaming kernel shared by KNN, three_nn, ball_query and Chamfer.
//
// Layout / algorithm (DESIGN.md "Neighbourhood engine"):
//   * a pack kernel converts the reference cloud once to SoA rows [B][ROWS][Npad] in the caller's
//     workspace (x, y, z and |r|^2 for the expanded form; negated coordinates for the direct form),
//     padded with sentinels that can never be selected;
//   * each CTA owns QT*CW*32 queries (held in registers) of one cloud and one split of the refs;
//     a producer warp streams 512-ref tiles of the SoA rows into a 3-stage shared-memory ring with
//     1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); consumer warps read a group of 4
//     refs with broadcast LDS.128 and evaluate them with packed FP32x2 math (FFMA2/FMUL2/FADD2);
//   * selection is a threshold filter: min over the group vs the query's current bound tau; a hit
//     only appends the group id to a per-query pending list (branch-free: one store + one
//     predicated pointer bump). Pending groups are re-evaluated bit-identically in warp-synchronous
//     drains and fed to the sink (bounded max-heap of (distance,index) keys / ball list), which
//     tightens tau. Refs are visited in ascending index order and keys order by (distance, index),
//     so the lowest index wins ties.
#pragma once
#include "common.cuh"

namespace b200pci {

constexpr int NBR_TILE = 512;   // refs per shared-memory stage
constexpr int NBR_STAGES = 3;   // TMA ring depth
constexpr int NBR_PEND = 32;    // pending group ids per query
constexpr int NBR_CHECK = 8;    // groups between pending-overflow checks

template <int MODE>
struct NbrRows {
    static constexpr int value = (MODE == B200PCI_DIST_EXPANDED) ? 4 : 3;
};

struct NbrParams {
    int S, N, Npad;
    int nsplit, tiles_per_split, total_tiles;
    const float *q;
    long long q_sb, q_sp, q_sc;
    const float *ws_ref;  // [B][ROWS][Npad]
};

// ---- pack kernel ---------------------------------------------------------------------------
template <int MODE>
__global__ void nbr_pack_refs_kernel(int N, int Npad, const float *__restrict__ r, long long r_sb,
                                     long long r_sp, long long r_sc, float *__restrict__ ws) {
    constexpr int ROWS = NbrRows<MODE>::value;
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    float *row = ws + (size_t)b * ROWS * Npad;
    const float inf = __int_as_float(0x7f800000);
    float x = 0.f, y = 0.f, z = 0.f, w = inf;
    if (j < N) {
        const float *p = r + b * r_sb + j * r_sp;
        x = p[0];
        y = p[r_sc];
        z = p[2 * r_sc];
        w = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    }
    if (MODE == B200PCI_DIST_EXPANDED) {
        row[j] = x;
        row[Npad + j] = y;
        row[2 * Npad + j] = z;
        row[3 * Npad + j] = w;
    } else {
        row[j] = (j < N) ? -x : -inf;  // d = q + (-r); padding -> (-inf)^2 = +inf
        row[Npad + j] = -y;
        row[2 * Npad + j] = -z;
    }
}

// ---- per-query constants and the 4-ref distance evaluation ------------------------------------
template <int MODE>
struct QueryRegs {
    f32x2 a, b, c, s;  // expanded: (-2x,-2x) (-2y,-2y) (-2z,-2z) (|q|^2,|q|^2); direct: (x,x) (y,y) (z,z)
    __device__ __forceinline__ void set(float x, float y, float z) {
        if (MODE == B200PCI_DIST_EXPANDED) {
            float sq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
            a = pack2(-2.f * x, -2.f * x);
            b = pack2(-2.f * y, -2.f * y);
            c = pack2(-2.f * z, -2.f * z);
            s = pack2(sq, sq);
        } else {
            a = pack2(x, x);
            b = pack2(y, y);
            c = pack2(z, z);
            s = 0;
        }
    }
};

// d[0..3] for refs (X.x..X.w, ...). Same instruction sequence in the main loop and in drains.
template <int MODE>
__device__ __forceinline__ void dist4(const QueryRegs<MODE> &q, const float4 &X, const float4 &Y,
                                      const float4 &Z, const float4 &W, float (&d)[4]) {
    if (MODE == B200PCI_DIST_EXPANDED) {
        // t = -2*dot = fma(-2z,Z,fma(-2y,Y,(-2x)*X)); D = (t + |q|^2) + |r|^2
        f32x2 t0 = mul2(pack2(X.x, X.y), q.a), t1 = mul2(pack2(X.z, X.w), q.a);
        t0 = fma2(pack2(Y.x, Y.y), q.b, t0);
        t1 = fma2(pack2(Y.z, Y.w), q.b, t1);
        t0 = fma2(pack2(Z.x, Z.y), q.c, t0);
        t1 = fma2(pack2(Z.z, Z.w), q.c, t1);
        t0 = add2(t0, q.s);
        t1 = add2(t1, q.s);
        t0 = add2(t0, pack2(W.x, W.y));
        t1 = add2(t1, pack2(W.z, W.w));
        unpack2(t0, d[0], d[1]);
        unpack2(t1, d[2], d[3]);
    } else {
        // rows hold -r: dx = q + (-X); D = fma(dz,dz,fma(dx,dx,dy*dy))
        f32x2 x0 = add2(pack2(X.x, X.y), q.a), x1 = add2(pack2(X.z, X.w), q.a);
        f32x2 y0 = add2(pack2(Y.x, Y.y), q.b), y1 = add2(pack2(Y.z, Y.w), q.b);
        f32x2 z0 = add2(pack2(Z.x, Z.y), q.c), z1 = add2(pack2(Z.z, Z.w), q.c);
        f32x2 t0 = mul2(y0, y0), t1 = mul2(y1, y1);
        t0 = fma2(x0, x0, t0);
        t1 = fma2(x1, x1, t1);
        t0 = fma2(z0, z0, t0);
        t1 = fma2(z1, z1, t1);
        unpack2(t0, d[0], d[1]);
        unpack2(t1, d[2], d[3]);
    }
}

// ---- sinks -----------------------------------------------------------------------------------
// A sink owns the per-query selection state. Interface (all called warp-synchronously):
//   smem_bytes(nt)           shared memory needed for nt query slots
//   init(...)                per-thread setup, returns initial tau for query slot j
//   offer(j, act, d, idx)    candidate from a drain
//   tau(j)                   current bound (hit test is d < tau)
//   finish(j, ...)           write results

// Bounded max-heap of 64-bit keys (sortable(distance) << 32 | index): the K smallest keys.
template <int K, int QT, int NT>
struct TopKSink {
    struct Params {
        void *idx;             // final: int64/int32 [B,S,K]   (nsplit == 1)
        float *dist;           // final, nullable
        int idx_is_int64;
        unsigned long long *part;  // partial keys [B,S,nsplit,K] (nsplit > 1)
    };
    static constexpr int LEVELS = (K >= 64) ? 6 : (K >= 32) ? 5 : (K >= 16) ? 4 : (K >= 8) ? 3
                                  : (K >= 4) ? 2 : (K >= 2) ? 1 : 0;
    static __host__ __device__ constexpr size_t smem_bytes() {
        return (K > 1) ? (size_t)K * QT * NT * sizeof(unsigned long long) : 0;
    }
    unsigned long long *heap;  // [QT][K][NT], this thread's column
    unsigned long long root[QT];

    __device__ __forceinline__ unsigned long long &H(int j, int node) {
        return heap[(size_t)(j * K + node) * NT];
    }
    __device__ __forceinline__ void init(unsigned char *smem, int tid) {
        heap = reinterpret_cast<unsigned long long *>(smem) + tid;
#pragma unroll
        for (int j = 0; j < QT; ++j) {
            root[j] = B200PCI_KEY_INF;
            if (K > 1)
                for (int n = 0; n < K; ++n) H(j, n) = B200PCI_KEY_INF;
        }
    }
    __device__ __forceinline__ float tau(int j) const {
        return sortable2f((uint32_t)(root[j] >> 32));
    }
    // replace the root by `key` (key < root for lanes with h) and restore the heap property
    __device__ __forceinline__ void replace_root(int j, bool h, unsigned long long key) {
        if (K == 1) {
            if (h) root[j] = key;
            return;
        }
        int pos = 0;
        bool moving = h;
#pragma unroll
        for (int l = 0; l < LEVELS; ++l) {
            const int c1 = 2 * pos + 1, c2 = c1 + 1;
            unsigned long long k1 = (c1 < K) ? H(j, c1) : 0ull;
            unsigned long long k2 = (c2 < K) ? H(j, c2) : 0ull;
            const bool right = k2 > k1;
            const unsigned long long kb = right ? k2 : k1;
            const bool down = moving && (kb > key);
            if (down) {
                H(j, pos) = kb;
                pos = right ? c2 : c1;
            } else if (moving) {
                H(j, pos) = key;
                moving = false;
            }
        }
        if (moving) H(j, pos) = key;
        root[j] = H(j, 0);
    }
    __device__ __forceinline__ void offer(int j, bool act, float d, uint32_t idx) {
        const unsigned long long key = make_key(d, idx);
        const bool h = act && (key < root[j]);
        if (__any_sync(0xffffffffu, h)) replace_root(j, h, key);
    }
    // qidx: query index within the cloud (or -1 if this slot has no query)
    // kout <= K: number of neighbours the caller asked for.
    __device__ __forceinline__ void finish(int j, const Params &p, int b, int S, int qidx,
                                           int nsplit, int split, int kout) {
        const bool valid = qidx >= 0;
        const size_t qrow = (size_t)b * S + (valid ? qidx : 0);
#pragma unroll 1
        for (int i = K - 1; i >= 0; --i) {
            const unsigned long long top = root[j];
            if (valid && i < kout) {
                if (nsplit > 1) {
                    p.part[(qrow * nsplit + split) * kout + i] = top;
                } else {
                    const uint32_t id = (uint32_t)top;
                    if (p.idx_is_int64)
                        reinterpret_cast<long long *>(p.idx)[qrow * kout + i] = (long long)id;
                    else
                        reinterpret_cast<int *>(p.idx)[qrow * kout + i] = (int)id;
                    if (p.dist) p.dist[qrow * kout + i] = sortable2f((uint32_t)(top >> 32));
                }
            }
            if (K > 1) replace_root(j, true, 0ull);  // pop: a minimal key sinks to a leaf
        }
    }
};

// ball_query: first `nsample` indices (ascending) with d < r^2, remaining slots = first hit.
// pointnet2/src/ball_query_gpu.cu:30-44.
template <int QT, int NT>
struct BallSink {
    struct Params {
        int *idx;  // [B,S,nsample], pre-zeroed by the caller
        int nsample;
        float radius2;
    };
    static __host__ __device__ constexpr size_t smem_bytes() { return 0; }
    int cnt[QT];
    int *row[QT];
    float r2;
    int ns;
    __device__ __forceinline__ void init(unsigned char *, int) {}
    __device__ __forceinline__ void setup(const Params &p, int j, int b, int S, int qidx) {
        r2 = p.radius2;
        ns = p.nsample;
        cnt[j] = (qidx >= 0) ? 0 : p.nsample;  // slots without a query are "full"
        row[j] = p.idx + ((size_t)b * S + (qidx >= 0 ? qidx : 0)) * p.nsample;
    }
    __device__ __forceinline__ float tau(int j) const {
        return (cnt[j] < ns) ? r2 : __int_as_float(0xff800000);  // -inf: never hit again
    }
    __device__ __forceinline__ void offer(int j, bool act, float d, uint32_t idx) {
        if (act && d < r2 && cnt[j] < ns) {
            if (cnt[j] == 0)
                for (int l = 0; l < ns; ++l) row[j][l] = (int)idx;
            row[j][cnt[j]] = (int)idx;
            ++cnt[j];
        }
    }
};

// ---- the streaming kernel --------------------------------------------------------------------
template <int MODE, int QT, int CW>
struct NbrSmem {
    static constexpr int ROWS = NbrRows<MODE>::value;
    static constexpr int NT = CW * 32;
    static constexpr size_t tiles_bytes = (size_t)NBR_STAGES * ROWS * NBR_TILE * sizeof(float);
    static constexpr size_t bars_bytes = 2 * NBR_STAGES * sizeof(uint64_t);
    static constexpr size_t pend_bytes = (size_t)QT * NBR_PEND * NT * sizeof(uint32_t);
    static constexpr size_t sink_off = tiles_bytes + 64 /*bars, padded*/ + pend_bytes;
};

// Returns after the last drain. `setup(sink, j, b, qidx)` runs once per query slot before the
// scan, `finish(sink, j, b, qidx, split)` once after it (consumer threads only).
template <int MODE, int QT, int CW, class Sink, class Setup, class Finish>
__device__ __forceinline__ void nbr_stream(const NbrParams &p, Sink &sink, Setup &&setup,
                                           Finish &&finish) {
    using SM = NbrSmem<MODE, QT, CW>;
    constexpr int ROWS = SM::ROWS;
    constexpr int NT = SM::NT;
    extern __shared__ __align__(128) unsigned char smem[];
    float *tiles = reinterpret_cast<float *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::tiles_bytes);
    uint64_t *empty = full + NBR_STAGES;
    uint32_t *pend = reinterpret_cast<uint32_t *>(smem + SM::tiles_bytes + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, split = blockIdx.y;
    const int tile0 = split * p.tiles_per_split;
    const int ntiles = min(p.tiles_per_split, p.total_tiles - tile0);
    const float *ws = p.ws_ref + (size_t)b * ROWS * p.Npad;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBR_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == CW) {  // ---- producer warp: one lane drives the TMA ring ----
        if (lane == 0) {
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % NBR_STAGES;
                if (t >= NBR_STAGES) mbar_wait(&empty[s], ((t / NBR_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&full[s], ROWS * NBR_TILE * sizeof(float));
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
                    tma_load_1d(tiles + (size_t)(s * ROWS + r) * NBR_TILE,
                                ws + (size_t)r * p.Npad + (size_t)(tile0 + t) * NBR_TILE,
                                NBR_TILE * sizeof(float), &full[s]);
            }
        }
        return;
    }

    // ---- consumer warps ----
    const int tid = threadIdx.x;  // 0 .. NT-1
    QueryRegs<MODE> q[QT];
    float tau[QT];
    int qidx[QT];
    uint32_t *pbase[QT], *pp[QT];
    sink.init(smem + SM::sink_off, tid);
#pragma unroll
    for (int j = 0; j < QT; ++j) {
        const int qi = (blockIdx.x * QT + j) * NT + tid;
        qidx[j] = (qi < p.S) ? qi : -1;
        float x = 0.f, y = 0.f, z = 0.f;
        if (qi < p.S) {
            const float *src = p.q + b * p.q_sb + qi * p.q_sp;
            x = src[0];
            y = src[p.q_sc];
            z = src[2 * p.q_sc];
        }
        q[j].set(x, y, z);
        setup(sink, j, b, qidx[j]);
        tau[j] = (qi < p.S) ? sink.tau(j) : __int_as_float(0xff800000);
        pbase[j] = pend + (size_t)j * NBR_PEND * NT + tid;
        pp[j] = pbase[j];
        *pbase[j] = 0;
    }

    // re-evaluate the pending groups of query slot j and feed the sink
    auto drain = [&](int j) {
        const int n = (int)(pp[j] - pbase[j]) / NT;
        const int nmax = warp_max_i(n);
        for (int e = 0; e < nmax; ++e) {
            const bool act = e < n;
            const uint32_t gid = act ? pbase[j][(size_t)e * NT] : 0u;
            const float4 X = __ldg(reinterpret_cast<const float4 *>(ws) + gid);
            const float4 Y = __ldg(reinterpret_cast<const float4 *>(ws + p.Npad) + gid);
            const float4 Z = __ldg(reinterpret_cast<const float4 *>(ws + 2 * (size_t)p.Npad) + gid);
            float4 W = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ROWS == 4) W = __ldg(reinterpret_cast<const float4 *>(ws + 3 * (size_t)p.Npad) + gid);
            float d[4];
            dist4<MODE>(q[j], X, Y, Z, W, d);
#pragma unroll
            for (int i = 0; i < 4; ++i) sink.offer(j, act, d[i], gid * 4u + i);
        }
        pp[j] = pbase[j];
        if (qidx[j] >= 0) tau[j] = sink.tau(j);
    };

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % NBR_STAGES;
        mbar_wait(&full[s], (t / NBR_STAGES) & 1);
        const float4 *sX = reinterpret_cast<const float4 *>(tiles + (size_t)(s * ROWS) * NBR_TILE);
        const float4 *sY = sX + NBR_TILE / 4;
        const float4 *sZ = sY + NBR_TILE / 4;
        const float4 *sW = sZ + NBR_TILE / 4;
        uint32_t gid = (uint32_t)(tile0 + t) * (NBR_TILE / 4);
        for (int g0 = 0; g0 < NBR_TILE / 4; g0 += NBR_CHECK) {
#pragma unroll
            for (int u = 0; u < NBR_CHECK; ++u) {
                const float4 X = sX[g0 + u], Y = sY[g0 + u], Z = sZ[g0 + u];
                float4 W = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ROWS == 4) W = sW[g0 + u];
#pragma unroll
                for (int j = 0; j < QT; ++j) {
                    float d[4];
                    dist4<MODE>(q[j], X, Y, Z, W, d);
                    const float m = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));
                    *pp[j] = gid;
                    pp[j] += (m < tau[j]) ? NT : 0;
                }
                ++gid;
            }
            bool over = false;
#pragma unroll
            for (int j = 0; j < QT; ++j) over |= (pp[j] - pbase[j]) > (NBR_PEND - NBR_CHECK - 1) * NT;
            if (__any_sync(0xffffffffu, over)) {
#pragma unroll
                for (int j = 0; j < QT; ++j) drain(j);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int j = 0; j < QT; ++j) drain(j);
#pragma unroll
    for (int j = 0; j < QT; ++j) finish(sink, j, b, qidx[j], split);
}

}  // namespace b200pci
