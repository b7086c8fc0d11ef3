// Bandwidth-bound index ops: gather_points, group_points, three_interpolate (+ gradients).
// Replaces pointnet2/src/sampling_gpu.cu:8-83, group_points_gpu.cu:8-86,
// interpolate_gpu.cu:77-161.
//
// The reference launches one thread per output SCALAR with the channel on blockIdx.y, so the
// index (and weight) of every output point is re-read C times and each thread does an integer
// divide. Here a thread owns 4 consecutive outputs of the contiguous (npoint*nsample) axis, reads
// their indices once (one 128-bit load), and loops over a chunk of channels issuing 128-bit
// stores; the feature rows it gathers from ([N] floats each) stay L1/L2 resident.
#include <algorithm>

#include "common.cuh"

namespace b200pci {

#ifndef GTH_THREADS_V  // (developer variants: tools/variants.sh)
#define GTH_THREADS_V 256
#endif
#ifndef GTH_CCHUNK_V
#define GTH_CCHUNK_V 8
#endif
constexpr int GTH_THREADS = GTH_THREADS_V;
constexpr int GTH_CCHUNK = GTH_CCHUNK_V;  // channels per CTA (blockIdx.y)

// out[b,c,t] = points[b,c,idx[b,t]],  t in [0,T)  (T = npoints*nsample; gather: nsample = 1)
template <bool VEC>
__global__ void __launch_bounds__(GTH_THREADS)
    group_kernel(int C, int N, long long T, const float *__restrict__ points,
                 const int *__restrict__ idx, float *__restrict__ out, long long out_bs) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GTH_CCHUNK;
    const int c1 = min(C, c0 + GTH_CCHUNK);
    const long long t0 = ((long long)blockIdx.x * GTH_THREADS + threadIdx.x) * 4;
    if (t0 >= T) return;
    const int *ip = idx + (size_t)b * T + t0;
    int i0, i1 = 0, i2 = 0, i3 = 0;
    const bool full = VEC || (t0 + 3 < T);
    if (VEC) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(ip));
        i0 = v.x; i1 = v.y; i2 = v.z; i3 = v.w;
    } else {
        i0 = ip[0];
        if (t0 + 1 < T) i1 = ip[1];
        if (t0 + 2 < T) i2 = ip[2];
        if (t0 + 3 < T) i3 = ip[3];
    }
    const float *src = points + ((size_t)b * C + c0) * N;
    float *dst = out + (size_t)b * out_bs + (size_t)c0 * T + t0;  // out_bs = C * T unless the rows are a slice
#pragma unroll 4
    for (int c = c0; c < c1; ++c, src += N, dst += T) {
        const float v0 = __ldg(src + i0), v1 = __ldg(src + i1), v2 = __ldg(src + i2),
                    v3 = __ldg(src + i3);
        if (VEC) {
            __stcs(reinterpret_cast<float4 *>(dst), make_float4(v0, v1, v2, v3));
        } else if (full) {
            dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3;
        } else {
            dst[0] = v0;
            if (t0 + 1 < T) dst[1] = v1;
            if (t0 + 2 < T) dst[2] = v2;
        }
    }
}

// grad_points[b,c,idx[b,t]] += grad_out[b,c,t]   (atomic, like group_points_gpu.cu:24)
__global__ void __launch_bounds__(GTH_THREADS)
    group_grad_kernel(int C, int N, long long T, const float *__restrict__ grad_out,
                      const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GTH_CCHUNK;
    const int c1 = min(C, c0 + GTH_CCHUNK);
    const long long t = (long long)blockIdx.x * GTH_THREADS + threadIdx.x;
    if (t >= T) return;
    const int i = idx[(size_t)b * T + t];
    const float *g = grad_out + ((size_t)b * C + c0) * T + t;
    float *dst = grad_points + ((size_t)b * C + c0) * N + i;
    for (int c = c0; c < c1; ++c, g += T, dst += N) atomicAdd(dst, __ldg(g));
}

// out[b,c,i] = fma(w2,p2,fma(w0,p0,w1*p1)), p_j = points[b,c,idx[b,i,j]]
// (the contraction nvcc -O2 produces for interpolate_gpu.cu:96)
__global__ void __launch_bounds__(GTH_THREADS)
    three_interpolate_kernel(int C, int m, int n, const float *__restrict__ points,
                             const int *__restrict__ idx, const float *__restrict__ weight,
                             float *__restrict__ out) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GTH_CCHUNK;
    const int c1 = min(C, c0 + GTH_CCHUNK);
    const int i = blockIdx.x * GTH_THREADS + threadIdx.x;
    if (i >= n) return;
    const int *ip = idx + ((size_t)b * n + i) * 3;
    const float *wp = weight + ((size_t)b * n + i) * 3;
    const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
    const float w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const float *src = points + ((size_t)b * C + c0) * m;
    float *dst = out + ((size_t)b * C + c0) * n + i;
#pragma unroll 4
    for (int c = c0; c < c1; ++c, src += m, dst += n) {
        const float p0 = __ldg(src + i0), p1 = __ldg(src + i1), p2 = __ldg(src + i2);
        __stcs(dst, __fmaf_rn(w2, p2, __fmaf_rn(w0, p0, __fmul_rn(w1, p1))));
    }
}

// Shared-memory variant: the 3 random 4-byte gathers per output are the bottleneck (a warp-wide LDG
// that touches 32 different lines costs 32 L1 wavefronts). Here a CTA stages CC whole feature rows
// ([m] floats each, contiguous in [B,C,m]) in shared memory, where a 32-way random gather costs
// ~3.5 bank-conflict cycles, and sweeps a slice of the n output points over those rows.
#ifndef TI_THREADS_V
#define TI_THREADS_V 512
#endif
constexpr int TI_THREADS = TI_THREADS_V;
__global__ void __launch_bounds__(TI_THREADS)
    three_interpolate_rows_kernel(int C, int m, int n, int CC, int nslices,
                                  const float *__restrict__ points, const int *__restrict__ idx,
                                  const float *__restrict__ weight, float *__restrict__ out) {
    extern __shared__ __align__(16) float rows[];  // [CC][m]
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CC;
    const int cc = min(CC, C - c0);
    const float *src = points + ((size_t)b * C + c0) * m;
    const int total = cc * m;
    if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int t = threadIdx.x; t < total / 4; t += TI_THREADS)
            reinterpret_cast<float4 *>(rows)[t] = __ldg(reinterpret_cast<const float4 *>(src) + t);
    } else {
        for (int t = threadIdx.x; t < total; t += TI_THREADS) rows[t] = __ldg(src + t);
    }
    __syncthreads();
    const int per = (n + nslices - 1) / nslices;
    const int i_begin = blockIdx.x * per, i_end = min(n, i_begin + per);
    for (int i = i_begin + threadIdx.x; i < i_end; i += TI_THREADS) {
        const int *ip = idx + ((size_t)b * n + i) * 3;
        const float *wp = weight + ((size_t)b * n + i) * 3;
        const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
        const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
        float *dst = out + ((size_t)b * C + c0) * n + i;
        const float *r = rows;
#pragma unroll 4
        for (int c = 0; c < cc; ++c, r += m, dst += n)
            __stcs(dst, __fmaf_rn(w2, r[i2], __fmaf_rn(w0, r[i0], __fmul_rn(w1, r[i1]))));
    }
}

// Quad variant (the default for the model's / the bench's shapes). The rows variant above is bound by
// instruction issue, not by memory: ~10 instructions per output element (three 4-byte shared-memory
// gathers with ~3.5-way bank conflicts, three FP ops, a 4-byte store, address updates). Here a CTA
// stages FOUR channels interleaved per point (rows4[i] = {f[c0][i] .. f[c0+3][i]}: a neighbour's
// four channels are ONE conflict-light LDS.128) and a thread produces a 4 x 4 block -- four
// consecutive output points x four channels: 6 LDG.128 for its indices / weights, 12 LDS.128, the
// arithmetic on packed FP32x2 pairs (FMUL2 / FFMA2, IEEE per lane: the same bits as the scalar
// fma(w2,p2,fma(w0,p0,w1*p1))), and four 16-byte streaming stores along the contiguous point axis:
// ~3.5 instructions per output element instead of ~10.
constexpr int TQ_THREADS = 512;
static int g_ti_slices = 0;   // key 15 (developer): point slices of the quad kernel, 0 = automatic
static int g_ti_variant = 0;  // key 16 (developer): 1 = rows kernel instead of the quad kernel
__device__ __forceinline__ f32x2 ti_blend(f32x2 a, f32x2 b, f32x2 c, float w0, float w1, float w2) {
    return fma2(pack2(w2, w2), c, fma2(pack2(w0, w0), a, mul2(pack2(w1, w1), b)));
}
__global__ void __launch_bounds__(TQ_THREADS)
    three_interpolate_quad_kernel(int C, int m, int n, int nslices, const float *__restrict__ points,
                                  const int *__restrict__ idx, const float *__restrict__ weight,
                                  float *__restrict__ out) {
    extern __shared__ __align__(16) float4 rows4[];  // [m]
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 4;
    const int cc = min(4, C - c0);
    const float *src = points + ((size_t)b * C + c0) * m;
    if (cc == 4 && (m & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        // four points of each of the four channel rows per thread (4 LDG.128), transposed in registers
        const float4 *s0 = reinterpret_cast<const float4 *>(src), *s1 = s0 + m / 4, *s2 = s1 + m / 4,
                     *s3 = s2 + m / 4;
        for (int i = threadIdx.x; i < m / 4; i += TQ_THREADS) {
            const float4 a = __ldg(s0 + i), bq = __ldg(s1 + i), cq = __ldg(s2 + i), dq = __ldg(s3 + i);
            rows4[4 * i + 0] = make_float4(a.x, bq.x, cq.x, dq.x);
            rows4[4 * i + 1] = make_float4(a.y, bq.y, cq.y, dq.y);
            rows4[4 * i + 2] = make_float4(a.z, bq.z, cq.z, dq.z);
            rows4[4 * i + 3] = make_float4(a.w, bq.w, cq.w, dq.w);
        }
    } else {
        for (int i = threadIdx.x; i < m; i += TQ_THREADS)
            rows4[i] = make_float4(__ldg(src + i), cc > 1 ? __ldg(src + m + i) : 0.f,
                                   cc > 2 ? __ldg(src + 2 * (size_t)m + i) : 0.f,
                                   cc > 3 ? __ldg(src + 3 * (size_t)m + i) : 0.f);
    }
    __syncthreads();
    const int per = (((n + nslices - 1) / nslices) + 3) & ~3;  // slice length, a multiple of 4 (n % 4 == 0)
    const int i_begin = blockIdx.x * per, i_end = min(n, i_begin + per);
    float *dst = out + ((size_t)b * C + c0) * n;
    // the next block's indices / weights are fetched before the current block is gathered: a thread
    // has only 4-8 blocks, and with ~1 CTA per SM nothing else hides the L2 round trip
    const int4 *ibase = reinterpret_cast<const int4 *>(idx + (size_t)b * n * 3);
    const float4 *wbase = reinterpret_cast<const float4 *>(weight + (size_t)b * n * 3);
    int i = i_begin + threadIdx.x * 4;
    int4 ia, ib, ic;
    float4 wa, wb, wc;
    if (i < i_end) {
        const int4 *ip = ibase + (i / 4) * 3;
        const float4 *wp = wbase + (i / 4) * 3;
        ia = __ldg(ip), ib = __ldg(ip + 1), ic = __ldg(ip + 2);
        wa = __ldg(wp), wb = __ldg(wp + 1), wc = __ldg(wp + 2);
    }
    for (; i < i_end; i += TQ_THREADS * 4) {
        const int4 ca = ia, cbq = ib, cc3 = ic;
        const float4 va = wa, vb = wb, vc = wc;
        const int inext = i + TQ_THREADS * 4;
        if (inext < i_end) {
            const int4 *ip = ibase + (inext / 4) * 3;
            const float4 *wp = wbase + (inext / 4) * 3;
            ia = __ldg(ip), ib = __ldg(ip + 1), ic = __ldg(ip + 2);
            wa = __ldg(wp), wb = __ldg(wp + 1), wc = __ldg(wp + 2);
        }
        const int id[12] = {ca.x, ca.y, ca.z, ca.w, cbq.x, cbq.y, cbq.z, cbq.w, cc3.x, cc3.y, cc3.z, cc3.w};
        const float w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
        float o[4][4];  // [channel][point]
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float4 f0 = rows4[id[3 * p]], f1 = rows4[id[3 * p + 1]], f2 = rows4[id[3 * p + 2]];
            const f32x2 lo = ti_blend(pack2(f0.x, f0.y), pack2(f1.x, f1.y), pack2(f2.x, f2.y), w[3 * p],
                                      w[3 * p + 1], w[3 * p + 2]);
            const f32x2 hi = ti_blend(pack2(f0.z, f0.w), pack2(f1.z, f1.w), pack2(f2.z, f2.w), w[3 * p],
                                      w[3 * p + 1], w[3 * p + 2]);
            unpack2(lo, o[0][p], o[1][p]);
            unpack2(hi, o[2][p], o[3][p]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < cc)
                __stcs(reinterpret_cast<float4 *>(dst + (size_t)c * n + i),
                       make_float4(o[c][0], o[c][1], o[c][2], o[c][3]));
    }
}

// interpolate_gpu.cu:139-141
__global__ void __launch_bounds__(GTH_THREADS)
    three_interpolate_grad_kernel(int C, int n, int m, const float *__restrict__ grad_out,
                                  const int *__restrict__ idx, const float *__restrict__ weight,
                                  float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GTH_CCHUNK;
    const int c1 = min(C, c0 + GTH_CCHUNK);
    const int i = blockIdx.x * GTH_THREADS + threadIdx.x;
    if (i >= n) return;
    const int *ip = idx + ((size_t)b * n + i) * 3;
    const float *wp = weight + ((size_t)b * n + i) * 3;
    const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
    const float w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const float *g = grad_out + ((size_t)b * C + c0) * n + i;
    float *dst = grad_points + ((size_t)b * C + c0) * m;
    for (int c = c0; c < c1; ++c, g += n, dst += m) {
        const float gv = __ldg(g);
        atomicAdd(dst + i0, __fmul_rn(gv, w0));
        atomicAdd(dst + i1, __fmul_rn(gv, w1));
        atomicAdd(dst + i2, __fmul_rn(gv, w2));
    }
}

static int check_grid(int b, int c, const char *what) {
    B200PCI_CHECK_ARG(b <= 65535, "%s: batch %d exceeds grid limit", what, b);
    B200PCI_CHECK_ARG(ceil_div(c, GTH_CCHUNK) <= 65535, "%s: too many channels", what);
    return 0;
}

static int group_impl(int b, int c, int n, long long T, const float *points, const int *idx,
                      float *out, cudaStream_t st, const char *what, long long out_bs = -1) {
    if (out_bs < 0) out_bs = (long long)c * T;
    B200PCI_CHECK_ARG(b >= 0 && c >= 0 && n >= 0 && T >= 0, "%s: negative size", what);
    if (b == 0 || c == 0 || T == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(points && idx && out, "%s: null pointer", what);
    if (int rc = check_grid(b, c, what)) return rc;
    const bool vec = (T % 4 == 0) && (out_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    dim3 grid((unsigned)((T + 4LL * GTH_THREADS - 1) / (4LL * GTH_THREADS)), ceil_div(c, GTH_CCHUNK), b);
    if (vec)
        group_kernel<true><<<grid, GTH_THREADS, 0, st>>>(c, n, T, points, idx, out, out_bs);
    else
        group_kernel<false><<<grid, GTH_THREADS, 0, st>>>(c, n, T, points, idx, out, out_bs);
    B200PCI_LAUNCH_CHECK(what);
    return B200PCI_OK;
}

static int group_grad_impl(int b, int c, int n, long long T, const float *grad_out, const int *idx,
                           float *grad_points, cudaStream_t st, const char *what) {
    B200PCI_CHECK_ARG(b >= 0 && c >= 0 && n >= 0 && T >= 0, "%s: negative size", what);
    if (b == 0 || c == 0 || T == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(grad_out && idx && grad_points, "%s: null pointer", what);
    if (int rc = check_grid(b, c, what)) return rc;
    dim3 grid((unsigned)((T + GTH_THREADS - 1) / GTH_THREADS), ceil_div(c, GTH_CCHUNK), b);
    group_grad_kernel<<<grid, GTH_THREADS, 0, st>>>(c, n, T, grad_out, idx, grad_points);
    B200PCI_LAUNCH_CHECK(what);
    return B200PCI_OK;
}

// ---- row gathers in the [B,N,C] layout of models/pointconv_util.py --------------------------------
// out[b,t,:] = points[b, idx[b,t], :],  t in [0,T)  (T = S*K for index_points_group, S for
// index_points_gather). `points` may be any strided view (the model passes permuted [B,C,N]
// tensors); idx is int64 (what knn_point returns) or int32; out is contiguous [B,T,C].
// A thread moves one 4-channel piece of one row: consecutive threads write consecutive 16-byte
// pieces of `out` (the only HBM stream that matters: rows are re-read from L2).
// (32-bit piece index, division by the pieces-per-row constant as a multiply-high: the 64-bit
// divide of the first version made the kernel instruction-bound -- 56 % issue utilisation at
// 3.4 TB/s, profiles/r2_all_ops_ncu_summary.txt)
// Each thread moves RG_UNROLL independent 16-byte pieces (index loads first, then the row loads,
// then the stores): one piece per thread leaves 32 KB in flight per SM, which at the ~1.2 us of an
// index load followed by a dependent row load caps the kernel near 3.9 TB/s (measured 3.65-4.0).
constexpr int RG_UNROLL = 4;
// VEC_IN: unit channel stride and 16-byte aligned rows (one LDG.128 per piece); otherwise four
// scalar loads with the channel stride (channel-major [B,C,N] tables seen through a permuted view:
// every float is its own sector, like group_points). VEC_OUT: C % 4 == 0 and an aligned output.
template <bool VEC_IN, bool VEC_OUT>
__global__ void __launch_bounds__(GTH_THREADS)
    rows_gather_kernel(int N, long long T, int C, const float *__restrict__ points, long long p_sb,
                       long long p_sn, long long p_sc, const void *__restrict__ idx,
                       int idx_is_int64, float *__restrict__ out, FastDiv fcq) {
    const uint32_t cq = fcq.d;  // pieces per row
    const uint32_t total = (uint32_t)T * cq;
    // per-cloud bases once (the only 64-bit multiplies left are i * p_sn and t * C)
    const int b = blockIdx.y;
    const float *pts = points + b * p_sb;
    float *ob = out + (size_t)b * T * C;
    const uint32_t g0 = blockIdx.x * (GTH_THREADS * RG_UNROLL) + threadIdx.x;
    uint32_t t[RG_UNROLL], c0[RG_UNROLL];
    long long i[RG_UNROLL];
#pragma unroll
    for (int u = 0; u < RG_UNROLL; ++u) {
        const uint32_t g = g0 + u * GTH_THREADS;
        t[u] = fastdiv(min(g, total - 1), fcq);
        c0[u] = (min(g, total - 1) - t[u] * cq) * 4u;
        i[u] = idx_is_int64 ? reinterpret_cast<const long long *>(idx)[(size_t)b * T + t[u]]
                            : (long long)reinterpret_cast<const int *>(idx)[(size_t)b * T + t[u]];
    }
    float4 v[RG_UNROLL];
#pragma unroll
    for (int u = 0; u < RG_UNROLL; ++u) {
        if (VEC_IN) {
            v[u] = __ldg(reinterpret_cast<const float4 *>(pts + i[u] * p_sn + c0[u]));
        } else {
            const float *src = pts + i[u] * p_sn + (long long)c0[u] * p_sc;
            const uint32_t left = (uint32_t)C - c0[u];  // >= 1
            v[u].x = __ldg(src);
            v[u].y = left > 1 ? __ldg(src + p_sc) : 0.f;
            v[u].z = left > 2 ? __ldg(src + 2 * p_sc) : 0.f;
            v[u].w = left > 3 ? __ldg(src + 3 * p_sc) : 0.f;
        }
    }
#pragma unroll
    for (int u = 0; u < RG_UNROLL; ++u) {
        if (g0 + u * GTH_THREADS >= total) continue;
        float *dst = ob + (size_t)t[u] * C + c0[u];
        if (VEC_OUT) {
            __stcs(reinterpret_cast<float4 *>(dst), v[u]);
        } else {
            const uint32_t left = (uint32_t)C - c0[u];
            dst[0] = v[u].x;
            if (left > 1) dst[1] = v[u].y;
            if (left > 2) dst[2] = v[u].z;
            if (left > 3) dst[3] = v[u].w;
        }
    }
}

// grad_points[b, idx[b,t], :] += grad_out[b,t,:]  (atomic; grad_points contiguous [B,N,C], pre-zeroed)
__global__ void __launch_bounds__(GTH_THREADS)
    rows_gather_grad_kernel(int N, long long T, int C, const float *__restrict__ grad_out,
                            const void *__restrict__ idx, int idx_is_int64,
                            float *__restrict__ grad_points) {
    const long long g = (long long)blockIdx.x * GTH_THREADS + threadIdx.x;
    const int b = blockIdx.y;
    if (g >= T * C) return;
    const long long t = g / C;
    const int c = (int)(g - t * C);
    const long long i = idx_is_int64 ? reinterpret_cast<const long long *>(idx)[(size_t)b * T + t]
                                     : (long long)reinterpret_cast<const int *>(idx)[(size_t)b * T + t];
    atomicAdd(grad_points + ((size_t)b * N + i) * C + c, __ldg(grad_out + ((size_t)b * T + t) * C + c));
}

// ---- fused group / group_query of models/pointconv_util.py:194-241 ------------------------------------
// out[b,s,k,:] = [ xyz[b,idx[b,s,k],:] - centre[b,s,:]  |  points[b,idx[b,s,k],:] ]   (3 + D floats)
// norm[b,s,k,:] = the first three (the reference returns them as a second tensor).
// One thread per output float, consecutive threads -> consecutive addresses; the index of a row is
// loaded by all threads of the row (one L1 transaction). Replaces two gathers, a broadcast
// subtraction and a concatenation: five passes over the grouped tensor in the reference.
// A thread writes FOUR consecutive floats of `out` (one 16-byte streaming store; the four may straddle
// two rows) with 32-bit index arithmetic -- the flat index is split into (row, column) by
// multiply-high divisions by the run-time constants W = 3 + D and K. (The first version, one thread
// per float with 64-bit divides, ran at 0.6 TB/s with 72 % of its issue slots busy.)
template <bool VEC>
__global__ void __launch_bounds__(GTH_THREADS)
    group_concat_kernel(int N, int S, int K, int D, const float *__restrict__ xyz, long long x_sb, long long x_sn,
                        long long x_sc, const float *__restrict__ centre, long long c_sb, long long c_sn,
                        long long c_sc, const float *__restrict__ points, long long p_sb, long long p_sn,
                        long long p_sc, const void *__restrict__ idx, int idx_is_int64, float *__restrict__ out,
                        float *__restrict__ norm, FastDiv fW, FastDiv fK, uint32_t per_cloud) {
    const uint32_t W = fW.d;
    const uint32_t g0 = (blockIdx.x * GTH_THREADS + threadIdx.x) * 4u;
    const int b = blockIdx.y;
    if (g0 >= per_cloud) return;
    uint32_t t = fastdiv(g0, fW);  // (s, k) row
    uint32_t c = g0 - t * W;
    // per-cloud bases once; per ROW (at most two per thread) the three row pointers: the 64-bit
    // multiplies of the first version (3 per element) made the kernel instruction-bound
    const float *xb = xyz + b * x_sb, *cb = centre + b * c_sb, *pb = points ? points + b * p_sb : nullptr;
    const size_t row0 = (size_t)b * S * K;
    const float *xr, *cr, *pr;
    auto load_row = [&](uint32_t tt) {
        const long long i = idx_is_int64 ? reinterpret_cast<const long long *>(idx)[row0 + tt]
                                         : (long long)reinterpret_cast<const int *>(idx)[row0 + tt];
        xr = xb + i * x_sn;
        cr = cb + (long long)fastdiv(tt, fK) * c_sn;
        pr = pb + i * p_sn - 3 * p_sc;  // (column c of the row reads pr[c * p_sc], c >= 3)
    };
    load_row(t);
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        v[u] = 0.f;
        if (g0 + u < per_cloud) {
            if (c < 3u) {
                v[u] = __fsub_rn(__ldg(xr + (long long)c * x_sc), __ldg(cr + (long long)c * c_sc));
                if (norm != nullptr) norm[(row0 + t) * 3 + c] = v[u];
            } else {
                v[u] = __ldg(pr + (long long)c * p_sc);
            }
            if (++c == W) {
                c = 0u;
                ++t;
                if (u < 3 && g0 + u + 1 < per_cloud) load_row(t);
            }
        }
    }
    if (out == nullptr) return;
    float *dst = out + (size_t)b * per_cloud + g0;
    if (VEC) {
        __stcs(reinterpret_cast<float4 *>(dst), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (g0 + u < per_cloud) dst[u] = v[u];
    }
}

// The same through shared memory: a CTA assembles GC_ROWS consecutive (s, k) rows -- one contiguous
// span of the output -- in shared memory and writes the span with aligned 16-byte stores. The
// gathers read whole 16-byte pieces of a feature row (D % 4 == 0, unit channel stride) or single
// floats; the per-element work is a piece decode (multiply-high division) and one load, instead of
// the ~100 instructions per element of the thread-per-4-outputs kernel above, whose threads
// straddle rows of odd width W = 3 + D and diverge on the xyz / feature boundary (ncu: 82 % issue
// utilisation at 1.1 TB/s). grid (ceil(S*K / rows), B), dynamic shared memory rows * W floats.
constexpr int GC_THREADS = 256;
template <bool VEC_IN>
__global__ void __launch_bounds__(GC_THREADS)
    group_concat_rows_kernel(int rows_per_cta, int SK, int D, const float *__restrict__ xyz, long long x_sb,
                             long long x_sn, long long x_sc, const float *__restrict__ centre, long long c_sb,
                             long long c_sn, long long c_sc, const float *__restrict__ points, long long p_sb,
                             long long p_sn, long long p_sc, const void *__restrict__ idx, int idx_is_int64,
                             float *__restrict__ out, float *__restrict__ norm, FastDiv fK, FastDiv fPieces,
                             int vec_out) {
    extern __shared__ __align__(16) float stage[];  // [rows][W]
    __shared__ long long srow[64];                  // gathered point of each row
    const int W = 3 + D;
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_cta;
    const int rows = min(rows_per_cta, SK - r0);
    const size_t row_base = (size_t)b * SK + r0;
    if ((int)threadIdx.x < rows)
        srow[threadIdx.x] = idx_is_int64 ? reinterpret_cast<const long long *>(idx)[row_base + threadIdx.x]
                                         : (long long)reinterpret_cast<const int *>(idx)[row_base + threadIdx.x];
    __syncthreads();
    // relative coordinates: columns 0..2
    if ((int)threadIdx.x < rows * 3) {
        const int r = threadIdx.x / 3, c = threadIdx.x - r * 3;
        const uint32_t sq = fastdiv((uint32_t)(r0 + r), fK);  // centre of the row
        const float v = __fsub_rn(__ldg(xyz + b * x_sb + srow[r] * x_sn + c * x_sc),
                                  __ldg(centre + b * c_sb + (long long)sq * c_sn + c * c_sc));
        stage[r * W + c] = v;
        if (norm != nullptr) norm[(row_base + r) * 3 + c] = v;
    }
    // features: columns 3..W-1
    const float *pb = points + b * p_sb;
    if (VEC_IN) {
        const uint32_t pieces = fPieces.d;  // D / 4
        for (uint32_t e = threadIdx.x; e < (uint32_t)rows * pieces; e += GC_THREADS) {
            const uint32_t r = fastdiv(e, fPieces), j = e - r * pieces;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(pb + srow[r] * p_sn) + j);
            float *d = stage + r * W + 3 + 4 * j;
            d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
        }
    } else if (D > 0) {
        const uint32_t cols = fPieces.d;  // D
        for (uint32_t e = threadIdx.x; e < (uint32_t)rows * cols; e += GC_THREADS) {
            const uint32_t r = fastdiv(e, fPieces), c = e - r * cols;
            stage[r * W + 3 + c] = __ldg(pb + srow[r] * p_sn + (long long)c * p_sc);
        }
    }
    __syncthreads();
    const int span = rows * W;
    float *dst = out + ((size_t)b * SK + r0) * W;
    if (vec_out && (span & 3) == 0) {
        for (int e = threadIdx.x; e < span / 4; e += GC_THREADS)
            __stcs(reinterpret_cast<float4 *>(dst) + e, reinterpret_cast<const float4 *>(stage)[e]);
    } else {
        for (int e = threadIdx.x; e < span; e += GC_THREADS) dst[e] = stage[e];
    }
}

// Rows of up to 320 floats (the model's: 35 ... 259): the same staging per WARP -- RW = 8 or 4
// consecutive rows, no CTA barrier, so the 64 resident warps of an SM overlap each other's
// index -> row -> store chains (the CTA-wide version above spends its time in three barrier-
// separated latency phases). grid (ceil(S*K / (8 RW)), B), dynamic shared memory 8 * RW * W floats.
template <bool VEC_IN, int RW>
__global__ void __launch_bounds__(GC_THREADS)
    group_concat_warp_kernel(int SK, int D, const float *__restrict__ xyz, long long x_sb, long long x_sn,
                             long long x_sc, const float *__restrict__ centre, long long c_sb, long long c_sn,
                             long long c_sc, const float *__restrict__ points, long long p_sb, long long p_sn,
                             long long p_sc, const void *__restrict__ idx, int idx_is_int64,
                             float *__restrict__ out, float *__restrict__ norm, FastDiv fK, FastDiv fPieces,
                             int vec_out) {
    extern __shared__ __align__(16) float stage_all[];
    const int W = 3 + D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *stage = stage_all + warp * (RW * W);
    const int b = blockIdx.y;
    const int r0 = (blockIdx.x * (GC_THREADS / 32) + warp) * RW;
    if (r0 >= SK) return;
    const int rows = min(RW, SK - r0);
    const size_t row_base = (size_t)b * SK + r0;
    long long mine = 0;  // lane r < rows: the gathered point of row r0 + r
    if (lane < rows)
        mine = idx_is_int64 ? reinterpret_cast<const long long *>(idx)[row_base + lane]
                            : (long long)reinterpret_cast<const int *>(idx)[row_base + lane];
    {   // relative coordinates: columns 0..2 (lanes 0 .. 3 rows - 1)
        const int r = lane / 3, c = lane - r * 3;
        const long long pi = __shfl_sync(0xffffffffu, mine, r < RW ? r : 0);
        if (r < rows) {
            const uint32_t sq = fastdiv((uint32_t)(r0 + r), fK);  // centre of the row
            const float v = __fsub_rn(__ldg(xyz + b * x_sb + pi * x_sn + c * x_sc),
                                      __ldg(centre + b * c_sb + (long long)sq * c_sn + c * c_sc));
            stage[r * W + c] = v;
            if (norm != nullptr) norm[(row_base + r) * 3 + c] = v;
        }
    }
    const float *pb = points + b * p_sb;
    const uint32_t per_row = fPieces.d;  // D / 4 (VEC_IN) or D
    const uint32_t n = (uint32_t)rows * per_row;
    if (D > 0) {
#pragma unroll 2
        for (uint32_t e0 = 0; e0 < n; e0 += 32) {
            const uint32_t e = min(e0 + lane, n - 1);
            const uint32_t r = fastdiv(e, fPieces), j = e - r * per_row;
            const long long pi = __shfl_sync(0xffffffffu, mine, (int)r);
            if (VEC_IN) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(pb + pi * p_sn) + j);
                if (e0 + lane < n) {
                    float *d = stage + r * W + 3 + 4 * j;
                    d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
                }
            } else {
                const float v = __ldg(pb + pi * p_sn + (long long)j * p_sc);
                if (e0 + lane < n) stage[r * W + 3 + j] = v;
            }
        }
    }
    __syncwarp();
    const int span = rows * W;
    float *dst = out + ((size_t)b * SK + r0) * W;
    if (vec_out && (span & 3) == 0 && ((RW * W) & 3) == 0) {
        for (int e = lane; e < span / 4; e += 32)
            __stcs(reinterpret_cast<float4 *>(dst) + e, reinterpret_cast<const float4 *>(stage)[e]);
    } else {
        for (int e = lane; e < span; e += 32) dst[e] = stage[e];
    }
}

// QueryAndGroup's coordinate part (pointnet2_utils.py:250-256: grouping_operation on the transposed
// cloud, then `-= new_xyz`): out[b, c, s, k] = xyz[b, idx[b,s,k], c] - new_xyz[b, s, c], c < 3, read from
// the row-major [B,N,3] cloud (no transposed copy), four neighbours per thread, written into the
// first three channel planes of a [B, Ctot, S*K] tensor (out_bs = Ctot * S * K).
template <bool VEC>
__global__ void __launch_bounds__(GTH_THREADS)
    group_xyz_rel_kernel(int N, int S, FastDiv fK, long long T, const float *__restrict__ xyz,
                         const float *__restrict__ centre, const int *__restrict__ idx,
                         float *__restrict__ out, long long out_bs) {
    const int b = blockIdx.y;
    const long long t0 = ((long long)blockIdx.x * GTH_THREADS + threadIdx.x) * 4;
    if (t0 >= T) return;
    const int *ip = idx + (size_t)b * T + t0;
    const float *xb = xyz + (size_t)b * N * 3, *cb = centre + (size_t)b * S * 3;
    float v[3][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const bool live = VEC || t0 + u < T;
        const int i = live ? ip[u] : 0;
        const uint32_t sq = fastdiv((uint32_t)(live ? t0 + u : t0), fK);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c][u] = __fsub_rn(__ldg(xb + (size_t)i * 3 + c), __ldg(cb + (size_t)sq * 3 + c));
    }
    float *dst = out + (size_t)b * out_bs + t0;
#pragma unroll
    for (int c = 0; c < 3; ++c, dst += T) {
        if (VEC) {
            __stcs(reinterpret_cast<float4 *>(dst), make_float4(v[c][0], v[c][1], v[c][2], v[c][3]));
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (t0 + u < T) dst[u] = v[c][u];
        }
    }
}

}  // namespace b200pci

using namespace b200pci;

extern "C" int b200pci_group_concat(int B, int N, int S, int K, int D, const float *xyz, int64_t x_sb, int64_t x_sn,
                                    int64_t x_sc, const float *centre, int64_t c_sb, int64_t c_sn, int64_t c_sc,
                                    const float *points, int64_t p_sb, int64_t p_sn, int64_t p_sc,
                                    const void *idx, int idx_is_int64, float *out, float *norm, void *stream) {
    B200PCI_CHECK_ARG(B >= 0 && N >= 0 && S >= 0 && K >= 0 && D >= 0, "group_concat: negative size");
    if (B == 0 || S == 0 || K == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(xyz && centre && idx && (out || norm), "group_concat: null pointer");
    B200PCI_CHECK_ARG(D == 0 || points, "group_concat: null points with D > 0");
    B200PCI_CHECK_ARG(B <= 65535, "group_concat: batch too large");
    if (out == nullptr) D = 0;  // only the relative coordinates are wanted
    const long long per_cloud = (long long)S * K * (3 + D);
    B200PCI_CHECK_ARG(per_cloud < (1LL << 31), "group_concat: more than 2^31 output floats per cloud");
    const FastDiv fW = make_fastdiv((uint32_t)(3 + D)), fK = make_fastdiv((uint32_t)K);
    const bool vec = per_cloud % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int W = 3 + D;
    const long long SK = (long long)S * K;
    int rows = std::min(64, (10240 / W) & ~3);  // rows per CTA: <= 40 KB of staging, a multiple of 4
    if (out != nullptr && rows >= 4) {
        const bool vec_in = D > 0 && D % 4 == 0 && p_sc == 1 && p_sn % 4 == 0 && p_sb % 4 == 0 &&
                            (reinterpret_cast<uintptr_t>(points) & 15) == 0;
        const FastDiv fP = make_fastdiv((uint32_t)(vec_in ? D / 4 : (D > 0 ? D : 1)));
        if (W <= 320) {  // per-warp staging
            const int rw = W <= 160 ? 8 : 4;
            const int per_cta = rw * (GC_THREADS / 32);
            dim3 wgrid((unsigned)((SK + per_cta - 1) / per_cta), B);
            const size_t wsmem = (size_t)per_cta * W * sizeof(float);
#define B200PCI_GC_WARP(V, R)                                                                                  \
    group_concat_warp_kernel<V, R><<<wgrid, GC_THREADS, wsmem, (cudaStream_t)stream>>>(                        \
        (int)SK, D, xyz, x_sb, x_sn, x_sc, centre, c_sb, c_sn, c_sc, points, p_sb, p_sn, p_sc, idx, idx_is_int64, \
        out, norm, fK, fP, vec ? 1 : 0)
            if (vec_in && rw == 8) B200PCI_GC_WARP(true, 8);
            else if (vec_in) B200PCI_GC_WARP(true, 4);
            else if (rw == 8) B200PCI_GC_WARP(false, 8);
            else B200PCI_GC_WARP(false, 4);
#undef B200PCI_GC_WARP
            B200PCI_LAUNCH_CHECK("group_concat_warp_kernel");
            return B200PCI_OK;
        }
        dim3 grid((unsigned)((SK + rows - 1) / rows), B);
        const size_t smem = (size_t)rows * W * sizeof(float);
        if (vec_in)
            group_concat_rows_kernel<true><<<grid, GC_THREADS, smem, (cudaStream_t)stream>>>(
                rows, (int)SK, D, xyz, x_sb, x_sn, x_sc, centre, c_sb, c_sn, c_sc, points, p_sb, p_sn, p_sc, idx,
                idx_is_int64, out, norm, fK, fP, vec ? 1 : 0);
        else
            group_concat_rows_kernel<false><<<grid, GC_THREADS, smem, (cudaStream_t)stream>>>(
                rows, (int)SK, D, xyz, x_sb, x_sn, x_sc, centre, c_sb, c_sn, c_sc, points, p_sb, p_sn, p_sc, idx,
                idx_is_int64, out, norm, fK, fP, vec ? 1 : 0);
        B200PCI_LAUNCH_CHECK("group_concat_rows_kernel");
        return B200PCI_OK;
    }
    // norm-only calls and rows wider than the staging area: one thread per 4 outputs
    dim3 grid((unsigned)((per_cloud + 4LL * GTH_THREADS - 1) / (4LL * GTH_THREADS)), B);
    if (vec)
        group_concat_kernel<true><<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(
            N, S, K, D, xyz, x_sb, x_sn, x_sc, centre, c_sb, c_sn, c_sc, points, p_sb, p_sn, p_sc, idx,
            idx_is_int64, out, norm, fW, fK, (uint32_t)per_cloud);
    else
        group_concat_kernel<false><<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(
            N, S, K, D, xyz, x_sb, x_sn, x_sc, centre, c_sb, c_sn, c_sc, points, p_sb, p_sn, p_sc, idx,
            idx_is_int64, out, norm, fW, fK, (uint32_t)per_cloud);
    B200PCI_LAUNCH_CHECK("group_concat_kernel");
    return B200PCI_OK;
}

extern "C" int b200pci_gather_points(int b, int c, int n, int npoints, const float *points,
                                     const int *idx, float *out, void *stream) {
    return group_impl(b, c, n, npoints, points, idx, out, (cudaStream_t)stream, "gather_points");
}
extern "C" int b200pci_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                                          const int *idx, float *grad_points, void *stream) {
    return group_grad_impl(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream,
                           "gather_points_grad");
}
extern "C" int b200pci_group_points(int b, int c, int n, int npoints, int nsample,
                                    const float *points, const int *idx, float *out, void *stream) {
    B200PCI_CHECK_ARG(npoints >= 0 && nsample >= 0, "group_points: negative size");
    return group_impl(b, c, n, (long long)npoints * nsample, points, idx, out, (cudaStream_t)stream,
                      "group_points");
}
extern "C" int b200pci_query_group(int b, int n, int npoints, int nsample, int c, const float *xyz,
                                   const float *new_xyz, const float *features, const int *idx, float *out,
                                   int use_xyz, void *stream) {
    B200PCI_CHECK_ARG(b >= 0 && n >= 0 && npoints >= 0 && nsample >= 0 && c >= 0, "query_group: negative size");
    B200PCI_CHECK_ARG(use_xyz || c > 0, "query_group: nothing to group");
    const long long T = (long long)npoints * nsample;
    if (b == 0 || T == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(idx && out && (!use_xyz || (xyz && new_xyz)) && (c == 0 || features), "query_group: null pointer");
    B200PCI_CHECK_ARG(b <= 65535 && T < (1LL << 31), "query_group: batch or group count too large");
    const int ctot = (use_xyz ? 3 : 0) + c;
    const long long out_bs = (long long)ctot * T;
    if (use_xyz) {
        const bool vec = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        const FastDiv fK = make_fastdiv((uint32_t)nsample);
        dim3 grid((unsigned)((T + 4LL * GTH_THREADS - 1) / (4LL * GTH_THREADS)), b);
        if (vec)
            group_xyz_rel_kernel<true><<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(n, npoints, fK, T, xyz, new_xyz,
                                                                                     idx, out, out_bs);
        else
            group_xyz_rel_kernel<false><<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(n, npoints, fK, T, xyz, new_xyz,
                                                                                      idx, out, out_bs);
        B200PCI_LAUNCH_CHECK("group_xyz_rel_kernel");
    }
    if (c > 0)
        return group_impl(b, c, n, T, features, idx, out + (use_xyz ? 3 * T : 0), (cudaStream_t)stream,
                          "query_group", out_bs);
    return B200PCI_OK;
}
extern "C" int b200pci_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                         const float *grad_out, const int *idx, float *grad_points,
                                         void *stream) {
    B200PCI_CHECK_ARG(npoints >= 0 && nsample >= 0, "group_points_grad: negative size");
    return group_grad_impl(b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points,
                           (cudaStream_t)stream, "group_points_grad");
}
extern "C" int b200pci_three_interpolate(int b, int c, int m, int n, const float *points,
                                         const int *idx, const float *weight, float *out,
                                         void *stream) {
    B200PCI_CHECK_ARG(b >= 0 && c >= 0 && n >= 0 && m >= 0, "three_interpolate: negative size");
    if (b == 0 || c == 0 || n == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(points && idx && weight && out, "three_interpolate: null pointer");
    if (int rc = check_grid(b, c, "three_interpolate")) return rc;
    // four interleaved channels per CTA in shared memory (quad kernel) when the points are a multiple
    // of four, the table fits and there is enough work
    const size_t quad_bytes = (size_t)m * sizeof(float4);
    if (g_ti_variant == 0 && m > 0 && n % 4 == 0 && quad_bytes <= 200 * 1024 && (long long)n * c >= 64 * 1024 &&
        (reinterpret_cast<uintptr_t>(idx) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int quads = ceil_div(c, 4);
        const int per_sm = (int)((220 * 1024) / (quad_bytes + 1024));
        const long long slots = (long long)sm_count() * (per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm));
        int nslices = 1;
        // Staging the table costs as much per entry as producing an output, so a CTA should sweep
        // many more points than the table holds: the points are sliced only while all CTAs still
        // fit one resident wave and a slice keeps at least 4 sweeps of the CTA.
        while ((long long)quads * b * nslices * 2 <= slots && n / (nslices * 2) >= 16 * TQ_THREADS) nslices *= 2;
        if (g_ti_slices > 0) nslices = g_ti_slices;
        auto kern = three_interpolate_quad_kernel;
        if (quad_bytes > 48 * 1024)
            B200PCI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)quad_bytes));
        dim3 grid(nslices, quads, b);
        kern<<<grid, TQ_THREADS, quad_bytes, (cudaStream_t)stream>>>(c, m, n, nslices, points, idx, weight, out);
        B200PCI_LAUNCH_CHECK("three_interpolate_quad_kernel");
        return B200PCI_OK;
    }
    // rows of up to 100 KB per CTA (two CTAs per SM) in shared memory when a row fits and there
    // is enough work
    const size_t row_bytes = (size_t)m * sizeof(float);
    const size_t budget = 100 * 1024;
    if (m > 0 && row_bytes <= budget && (long long)n * c >= 64 * 1024) {
        int CC = (int)(budget / row_bytes);
        if (CC > c) CC = c;
        // enough CTAs for two waves: shrink the channel chunk first, then slice the points
        const int sms = sm_count();
        while (CC > 4 && (long long)ceil_div(c, CC) * b < 2LL * sms) CC = (CC + 1) / 2;
        int nslices = 1;
        while ((long long)ceil_div(c, CC) * b * nslices < 2LL * sms && n / (nslices * 2) >= 4 * TI_THREADS)
            nslices *= 2;
        const size_t smem = (size_t)CC * row_bytes;
        auto kern = three_interpolate_rows_kernel;
        if (smem > 48 * 1024)
            B200PCI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(nslices, ceil_div(c, CC), b);
        kern<<<grid, TI_THREADS, smem, (cudaStream_t)stream>>>(c, m, n, CC, nslices, points, idx,
                                                              weight, out);
        B200PCI_LAUNCH_CHECK("three_interpolate_rows_kernel");
        return B200PCI_OK;
    }
    dim3 grid(ceil_div(n, GTH_THREADS), ceil_div(c, GTH_CCHUNK), b);
    three_interpolate_kernel<<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(c, m, n, points, idx,
                                                                            weight, out);
    B200PCI_LAUNCH_CHECK("three_interpolate_kernel");
    return B200PCI_OK;
}
extern "C" int b200pci_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                              const int *idx, const float *weight,
                                              float *grad_points, void *stream) {
    B200PCI_CHECK_ARG(b >= 0 && c >= 0 && n >= 0 && m >= 0, "three_interpolate_grad: negative size");
    if (b == 0 || c == 0 || n == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(grad_out && idx && weight && grad_points, "three_interpolate_grad: null pointer");
    if (int rc = check_grid(b, c, "three_interpolate_grad")) return rc;
    dim3 grid(ceil_div(n, GTH_THREADS), ceil_div(c, GTH_CCHUNK), b);
    three_interpolate_grad_kernel<<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(
        c, n, m, grad_out, idx, weight, grad_points);
    B200PCI_LAUNCH_CHECK("three_interpolate_grad_kernel");
    return B200PCI_OK;
}

// K3 / K4: index_points_group / index_points_gather of models/pointconv_util.py:168-192 without the
// transpose copy, the int64 -> int32 cast and the [B,C,S,K] intermediate of the reference.
extern "C" int b200pci_index_points_rows(int B, int N, long long T, int C, const float *points,
                                         int64_t p_sb, int64_t p_sn, int64_t p_sc, const void *idx,
                                         int idx_is_int64, float *out, void *stream) {
    B200PCI_CHECK_ARG(B >= 0 && N >= 0 && T >= 0 && C >= 0, "index_points_rows: negative size");
    if (B == 0 || T == 0 || C == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(points && idx && out, "index_points_rows: null pointer");
    B200PCI_CHECK_ARG(B <= 65535, "index_points_rows: batch too large");
    const long long pieces = T * ((C + 3) / 4);
    B200PCI_CHECK_ARG(pieces < (1LL << 31), "index_points_rows: more than 2^31 16-byte pieces per cloud");
    const bool vec_out = C % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const bool vec_in = vec_out && p_sc == 1 && p_sn % 4 == 0 && p_sb % 4 == 0 &&
                        (reinterpret_cast<uintptr_t>(points) & 15) == 0;
    const FastDiv fcq = make_fastdiv((uint32_t)((C + 3) / 4));
    dim3 grid((unsigned)((pieces + GTH_THREADS * RG_UNROLL - 1) / (GTH_THREADS * RG_UNROLL)), B);
    if (vec_in)
        rows_gather_kernel<true, true><<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(
            N, T, C, points, p_sb, p_sn, p_sc, idx, idx_is_int64, out, fcq);
    else if (vec_out)
        rows_gather_kernel<false, true><<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(
            N, T, C, points, p_sb, p_sn, p_sc, idx, idx_is_int64, out, fcq);
    else
        rows_gather_kernel<false, false><<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(
            N, T, C, points, p_sb, p_sn, p_sc, idx, idx_is_int64, out, fcq);
    B200PCI_LAUNCH_CHECK("rows_gather_kernel");
    return B200PCI_OK;
}

extern "C" int b200pci_index_points_rows_grad(int B, int N, long long T, int C,
                                              const float *grad_out, const void *idx,
                                              int idx_is_int64, float *grad_points, void *stream) {
    B200PCI_CHECK_ARG(B >= 0 && N >= 0 && T >= 0 && C >= 0, "index_points_rows_grad: negative size");
    if (B == 0 || T == 0 || C == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(grad_out && idx && grad_points, "index_points_rows_grad: null pointer");
    B200PCI_CHECK_ARG(B <= 65535, "index_points_rows_grad: batch too large");
    dim3 grid((unsigned)((T * C + GTH_THREADS - 1) / GTH_THREADS), B);
    rows_gather_grad_kernel<<<grid, GTH_THREADS, 0, (cudaStream_t)stream>>>(
        N, T, C, grad_out, idx, idx_is_int64, grad_points);
    B200PCI_LAUNCH_CHECK("rows_gather_grad_kernel");
    return B200PCI_OK;
}

// developer hooks of this file (b200pci_debug_set keys 15, 16)
int b200pci_gather_debug_set(int key, double value) {
    if (key == 15)
        g_ti_slices = (int)value;
    else if (key == 16)
        g_ti_variant = (int)value;
    else
        return B200PCI_EINVAL;
    return B200PCI_OK;
}
