// Approximate Earth-Mover distance -- replaces models/EMD/cuda/emd_kernel.cu.
//
// The reference runs each cloud pair on ONE CTA (<<<32,512>>>, grid-stride over the batch,
// emd_kernel.cu:42,192): 10 temperature levels x 3 full n x m sweeps by 512 threads. Here every
// sweep is its own grid-wide kernel with one THREAD PER ROW (xyz1 point for sweeps 1 and 3,
// xyz2 point for sweep 2), the other cloud streamed through shared memory in index order. A row's
// accumulation order is therefore exactly the reference's (sequential over the other cloud,
// same fused multiply-adds as its SASS), so `match` is reproduced bit for bit while the work
// spreads over n/128 CTAs per pair instead of one.
#include "common.cuh"

namespace b200pci {

#ifndef EMD_THREADS_V  // (developer variants: tools/variants.sh)
#define EMD_THREADS_V 64
#endif
#ifndef EMD_TILE_V
#define EMD_TILE_V 2048
#endif
#ifndef EMD_MB_V
#define EMD_MB_V 16
#endif
constexpr int EMD_THREADS = EMD_THREADS_V;
constexpr int EMD_TILE = EMD_TILE_V;  // points of the streamed cloud per shared-memory tile
constexpr int EMD_MB = EMD_MB_V;      // match values prefetched per step in sweep 3

__device__ __forceinline__ float emd_d(float ax, float ay, float az, float bx, float by, float bz) {
    // (b-a)^2 summed as the reference compiles it: FMUL dy*dy; FFMA dx*dx+.; FFMA dz*dz+.
    const float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay), dz = __fsub_rn(bz, az);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// exp(level * d). EXACT: the reference's `__expf(level * d)` as nvcc compiles it without -ftz (FMUL by
// log2(e), range test, MUFU.EX2 with the two predicated scalings that produce denormal results):
// five issue slots, needed where `match` is compared bit for bit. Otherwise (the tolerance-checked
// forward-only cost): `ex2.approx.ftz` of d * (level * log2 e), the product folded on the host --
// two slots. The folded product differs from the two-step one by <= 1 ulp of the argument, i.e.
// a relative 6e-8 * |arg| on terms of weight e^arg, and results below 2^-126 are dropped.
template <bool EXACT>
__device__ __forceinline__ float emd_exp(float level, float d) {
    if (EXACT) return __expf(__fmul_rn(level, d));
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(level, d)));
    return r;
}

__global__ void emd_init_kernel(int n, int m, float multiL, float multiR, float *temp) {
    // temp per pair: remainL[n] remainR[m] ratioL[n] ratioR[m]   (emd_kernel.cu:30)
    float *t = temp + (size_t)blockIdx.y * (n + m) * 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] = multiL;
    if (i < m) t[n + i] = multiR;
}

// SWEEP 1 (emd_kernel.cu:54-87): ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level*d)*remainR[l])
// SWEEP 3 (:125-158):            match[l][k] += w, remainL[k] -= sum_l w,  w = exp*ratioL[k]*ratioR[l]
// SWEEP 4 (forward-only emd_cost): sweep 3 without the match matrix -- cost = sum_{k,l} d*match is
//          linear in match, so every level's increment w is weighted with d and summed on the fly
//          (rowcost[k], FP64 across tiles and levels); remainL is updated with the same FP32 chain
//          as sweep 3, so the ratios of the later levels stay bit-identical to the reference's.
// RS = threads per row. RS = 1 (EMD_THREADS rows per CTA) keeps a row's sum one sequential chain,
// i.e. bit-identical to the reference -- the path behind approxmatch_forward, whose `match` is
// compared bitwise. With 16384 rows that is 110 threads per SM, far too few to hide the ex2 / FMA
// latencies, so the forward-only emd_cost (a tolerance-checked scalar) runs RS = 8: eight lanes
// stride over a row's tile and their partial sums are combined by a fixed shuffle tree
// (deterministic; reassociation moves the result by ~1e-7 relative).
template <int RS>
struct EmdGeom {
    static constexpr int threads = RS == 1 ? EMD_THREADS : 256;
    static constexpr int rows = threads / RS;
};
template <int RS>
__device__ __forceinline__ float emd_lane_sum(float v) {
#pragma unroll
    for (int o = RS / 2; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int SWEEP, int RS>
__global__ void __launch_bounds__(EmdGeom<RS>::threads)
    emd_rows1_kernel(int n, int m, float level, const float *__restrict__ xyz1,
                     const float *__restrict__ xyz2, float *__restrict__ match, float *temp,
                     double *__restrict__ rowcost) {
    constexpr int NT = EmdGeom<RS>::threads;
    static_assert(RS == 1 || SWEEP != 3, "the match-writing sweep is the bit-exact path");
    __shared__ float4 buf[EMD_TILE];
    const int b = blockIdx.y;
    xyz1 += (size_t)b * n * 3;
    xyz2 += (size_t)b * m * 3;
    float *t = temp + (size_t)b * (n + m) * 2;
    float *remainL = t, *remainR = t + n, *ratioL = t + n + m, *ratioR = t + n + m + n;
    float *mt = match + (size_t)b * n * m;
    const int k = blockIdx.x * EmdGeom<RS>::rows + threadIdx.x / RS;
    const int sub = threadIdx.x % RS;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f, rl = 0.f;
    if (k < n) {
        x1 = xyz1[k * 3 + 0];
        y1 = xyz1[k * 3 + 1];
        z1 = xyz1[k * 3 + 2];
        if (SWEEP >= 3) rl = ratioL[k];
    }
    double cost = 0.0;
    float suml = (SWEEP == 1 && sub == 0) ? 1e-9f : 0.f;
    const float *side = (SWEEP == 1) ? remainR : ratioR;
    for (int l0 = 0; l0 < m; l0 += EMD_TILE) {
        const int lend = min(m, l0 + EMD_TILE) - l0;
        for (int l = threadIdx.x; l < lend; l += NT)
            buf[l] = make_float4(xyz2[(l0 + l) * 3 + 0], xyz2[(l0 + l) * 3 + 1],
                                 xyz2[(l0 + l) * 3 + 2], side[l0 + l]);
        __syncthreads();
        if (k < n) {
            if (SWEEP == 1) {
#pragma unroll 8
                for (int l = sub; l < lend; l += RS) {
                    const float4 p = buf[l];
                    const float e = emd_exp<RS == 1>(level, emd_d(x1, y1, z1, p.x, p.y, p.z));
                    suml = __fmaf_rn(e, p.w, suml);
                }
            } else if (SWEEP == 4) {
                float csub = 0.f;  // one tile's worth in FP32, tiles and levels in FP64
#pragma unroll 8
                for (int l = sub; l < lend; l += RS) {
                    const float4 p = buf[l];
                    const float d = emd_d(x1, y1, z1, p.x, p.y, p.z);
                    const float tw = __fmul_rn(rl, emd_exp<RS == 1>(level, d));
                    csub = __fmaf_rn(d, __fmul_rn(tw, p.w), csub);
                    suml = __fmaf_rn(tw, p.w, suml);
                }
                cost += (double)csub;
            } else {
                // match[l][k] is a read-modify-write in global memory: fetch EMD_MB values ahead so
                // that their L2 latency overlaps instead of serialising the (ordered) row sum
                float *mp = mt + (size_t)l0 * n + k;
                int l = 0;
                for (; l + EMD_MB <= lend; l += EMD_MB) {
                    float mv[EMD_MB];
#pragma unroll
                    for (int u = 0; u < EMD_MB; ++u) mv[u] = mp[(size_t)(l + u) * n];
#pragma unroll
                    for (int u = 0; u < EMD_MB; ++u) {
                        const float4 p = buf[l + u];
                        const float e = __expf(__fmul_rn(level, emd_d(x1, y1, z1, p.x, p.y, p.z)));
                        const float tw = __fmul_rn(rl, e);
                        mp[(size_t)(l + u) * n] = __fmaf_rn(tw, p.w, mv[u]);
                        suml = __fmaf_rn(tw, p.w, suml);
                    }
                }
                for (; l < lend; ++l) {
                    const float4 p = buf[l];
                    const float e = __expf(__fmul_rn(level, emd_d(x1, y1, z1, p.x, p.y, p.z)));
                    const float tw = __fmul_rn(rl, e);
                    mp[(size_t)l * n] = __fmaf_rn(tw, p.w, mp[(size_t)l * n]);
                    suml = __fmaf_rn(tw, p.w, suml);
                }
            }
        }
        __syncthreads();
    }
    if (RS > 1) {  // combine the lanes of a row (all lanes of the warp take part)
        suml = emd_lane_sum<RS>(suml);
        if (SWEEP == 4) {
#pragma unroll
            for (int o = RS / 2; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
        }
    }
    if (k < n && sub == 0) {
        if (SWEEP == 1)
            ratioL[k] = __fdiv_rn(remainL[k], suml);
        else
            remainL[k] = fmaxf(0.0f, __fsub_rn(remainL[k], suml));
        if (SWEEP == 4) rowcost[(size_t)b * n + k] += cost;
    }
}

// SWEEP 2 (emd_kernel.cu:89-123): per xyz2 point l.
template <int RS>
__global__ void __launch_bounds__(EmdGeom<RS>::threads)
    emd_rows2_kernel(int n, int m, float level, const float *__restrict__ xyz1,
                     const float *__restrict__ xyz2, float *temp) {
    constexpr int NT = EmdGeom<RS>::threads;
    __shared__ float4 buf[EMD_TILE];
    const int b = blockIdx.y;
    xyz1 += (size_t)b * n * 3;
    xyz2 += (size_t)b * m * 3;
    float *t = temp + (size_t)b * (n + m) * 2;
    float *remainR = t + n, *ratioL = t + n + m, *ratioR = t + n + m + n;
    const int l = blockIdx.x * EmdGeom<RS>::rows + threadIdx.x / RS;
    const int sub = threadIdx.x % RS;
    float x2 = 0.f, y2 = 0.f, z2 = 0.f;
    if (l < m) {
        x2 = xyz2[l * 3 + 0];
        y2 = xyz2[l * 3 + 1];
        z2 = xyz2[l * 3 + 2];
    }
    float sumr = 0.f;
    for (int k0 = 0; k0 < n; k0 += EMD_TILE) {
        const int kend = min(n, k0 + EMD_TILE) - k0;
        for (int k = threadIdx.x; k < kend; k += NT)
            buf[k] = make_float4(xyz1[(k0 + k) * 3 + 0], xyz1[(k0 + k) * 3 + 1],
                                 xyz1[(k0 + k) * 3 + 2], ratioL[k0 + k]);
        __syncthreads();
        if (l < m) {
#pragma unroll 8
            for (int k = sub; k < kend; k += RS) {
                const float4 p = buf[k];
                const float e = emd_exp<RS == 1>(level, emd_d(p.x, p.y, p.z, x2, y2, z2));
                sumr = __fmaf_rn(e, p.w, sumr);
            }
        }
        __syncthreads();
    }
    if (RS > 1) sumr = emd_lane_sum<RS>(sumr);
    if (l < m && sub == 0) {
        const float rr = remainR[l];
        sumr = __fmul_rn(sumr, rr);
        const float consumption = fminf(__fdiv_rn(rr, __fadd_rn(sumr, 1e-9f)), 1.0f);
        ratioR[l] = __fmul_rn(consumption, rr);
        remainR[l] = fmaxf(0.0f, __fsub_rn(rr, sumr));
    }
}

// matchcost (emd_kernel.cu:204-247): rowsum[k] = sum_l d(k,l)*match[l][k] (sequential FMA chain
// like the reference's per-thread subsum), then an FP64 reduction of the row sums per pair.
__global__ void __launch_bounds__(EMD_THREADS)
    emd_cost_rows_kernel(int n, int m, const float *__restrict__ xyz1,
                         const float *__restrict__ xyz2, const float *__restrict__ match,
                         float *__restrict__ rowsum) {
    __shared__ float buf[EMD_TILE * 3];
    const int b = blockIdx.y;
    xyz1 += (size_t)b * n * 3;
    xyz2 += (size_t)b * m * 3;
    const float *mt = match + (size_t)b * n * m;
    const int k = blockIdx.x * EMD_THREADS + threadIdx.x;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (k < n) {
        x1 = xyz1[k * 3 + 0];
        y1 = xyz1[k * 3 + 1];
        z1 = xyz1[k * 3 + 2];
    }
    float sub = 0.f;
    for (int l0 = 0; l0 < m; l0 += EMD_TILE) {
        const int lend = min(m, l0 + EMD_TILE) - l0;
        for (int l = threadIdx.x; l < lend * 3; l += EMD_THREADS) buf[l] = xyz2[l0 * 3 + l];
        __syncthreads();
        if (k < n) {
#pragma unroll 4
            for (int l = 0; l < lend; ++l) {
                const float d = emd_d(x1, y1, z1, buf[l * 3 + 0], buf[l * 3 + 1], buf[l * 3 + 2]);
                sub = __fmaf_rn(d, __ldg(mt + (size_t)(l0 + l) * n + k), sub);
            }
        }
        __syncthreads();
    }
    if (k < n) rowsum[(size_t)b * n + k] = sub;
}

template <class T>
__global__ void emd_cost_reduce_kernel(int n, const T *__restrict__ rowsum, float *cost) {
    __shared__ double sh[32];
    const int b = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)rowsum[(size_t)b * n + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) cost[b] = (float)s;
    }
}

// matchcostgrad1 (emd_kernel.cu:337-359): thread per xyz1 point, sequential over xyz2 -- same order.
__global__ void __launch_bounds__(EMD_THREADS)
    emd_grad1_kernel(int n, int m, const float *__restrict__ grad_cost,
                     const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                     const float *__restrict__ match, float *__restrict__ grad1) {
    __shared__ float buf[EMD_TILE * 3];
    const int b = blockIdx.y;
    xyz1 += (size_t)b * n * 3;
    xyz2 += (size_t)b * m * 3;
    const float *mt = match + (size_t)b * n * m;
    const int l = blockIdx.x * EMD_THREADS + threadIdx.x;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (l < n) {
        x1 = xyz1[l * 3 + 0];
        y1 = xyz1[l * 3 + 1];
        z1 = xyz1[l * 3 + 2];
    }
    float dx = 0.f, dy = 0.f, dz = 0.f;
    for (int k0 = 0; k0 < m; k0 += EMD_TILE) {
        const int kend = min(m, k0 + EMD_TILE) - k0;
        for (int k = threadIdx.x; k < kend * 3; k += EMD_THREADS) buf[k] = xyz2[k0 * 3 + k];
        __syncthreads();
        if (l < n) {
#pragma unroll 4
            for (int k = 0; k < kend; ++k) {
                const float mv = __ldg(mt + (size_t)(k0 + k) * n + l);
                const float d = __fadd_rn(mv, mv);
                dx = __fmaf_rn(d, __fsub_rn(x1, buf[k * 3 + 0]), dx);
                dy = __fmaf_rn(d, __fsub_rn(y1, buf[k * 3 + 1]), dy);
                dz = __fmaf_rn(d, __fsub_rn(z1, buf[k * 3 + 2]), dz);
            }
        }
        __syncthreads();
    }
    if (l < n) {
        const float g = grad_cost[b];
        float *o = grad1 + ((size_t)b * n + l) * 3;
        o[0] = __fmul_rn(dx, g);
        o[1] = __fmul_rn(dy, g);
        o[2] = __fmul_rn(dz, g);
    }
}

// matchcostgrad2 (emd_kernel.cu:290-331): one warp per xyz2 point, coalesced over the match row.
__global__ void __launch_bounds__(256)
    emd_grad2_kernel(int n, int m, const float *__restrict__ grad_cost,
                     const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                     const float *__restrict__ match, float *__restrict__ grad2) {
    const int b = blockIdx.y;
    xyz1 += (size_t)b * n * 3;
    xyz2 += (size_t)b * m * 3;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= m) return;
    const float *mrow = match + (size_t)b * n * m + (size_t)k * n;
    const float x2 = xyz2[k * 3 + 0], y2 = xyz2[k * 3 + 1], z2 = xyz2[k * 3 + 2];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float mv = __ldg(mrow + j);
        const float d = __fadd_rn(mv, mv);
        sx = __fmaf_rn(__fsub_rn(x2, xyz1[j * 3 + 0]), d, sx);
        sy = __fmaf_rn(__fsub_rn(y2, xyz1[j * 3 + 1]), d, sy);
        sz = __fmaf_rn(__fsub_rn(z2, xyz1[j * 3 + 2]), d, sz);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_down_sync(0xffffffffu, sx, o);
        sy += __shfl_down_sync(0xffffffffu, sy, o);
        sz += __shfl_down_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) {
        const float g = grad_cost[b];
        float *o = grad2 + ((size_t)b * m + k) * 3;
        o[0] = __fmul_rn(sx, g);
        o[1] = __fmul_rn(sy, g);
        o[2] = __fmul_rn(sz, g);
    }
}

static size_t emd_temp_bytes(int B, int n, int m) {
    return align_up((size_t)(B > 0 ? B : 1) * ((size_t)n + m) * 2 * sizeof(float), 256);
}

}  // namespace b200pci

using namespace b200pci;

extern "C" size_t b200pci_emd_workspace_bytes(int B, int n, int m) {
    if (B < 0 || n < 0 || m < 0) return 256;
    // approxmatch temp, plus the row sums used by matchcost (float) / emd_cost (double)
    return emd_temp_bytes(B, n, m) + align_up((size_t)(B > 0 ? B : 1) * n * sizeof(double), 256);
}

extern "C" int b200pci_emd_approxmatch(int B, int n, int m, const float *xyz1, const float *xyz2,
                                       float *match, void *workspace, size_t workspace_bytes,
                                       void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200PCI_CHECK_ARG(B >= 0 && n >= 0 && m >= 0, "emd_approxmatch: negative size");
    if (B == 0 || n == 0 || m == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(xyz1 && xyz2 && match, "emd_approxmatch: null pointer");
    B200PCI_CHECK_ARG(B <= 65535, "emd_approxmatch: batch too large");
    if (!workspace || workspace_bytes < emd_temp_bytes(B, n, m)) {
        set_error("emd_approxmatch: workspace of %zu bytes required, got %zu",
                  emd_temp_bytes(B, n, m), workspace_bytes);
        return B200PCI_EWORKSPACE;
    }
    float *temp = reinterpret_cast<float *>(workspace);
    // emd_kernel.cu:33-38 (integer division)
    const float multiL = (n >= m) ? 1.f : (float)(m / n);
    const float multiR = (n >= m) ? (float)(n / m) : 1.f;
    B200PCI_CUDA(cudaMemsetAsync(match, 0, (size_t)B * n * m * sizeof(float), st));
    const int mx = n > m ? n : m;
    emd_init_kernel<<<dim3(ceil_div(mx, 256), B), 256, 0, st>>>(n, m, multiL, multiR, temp);
    B200PCI_LAUNCH_CHECK("emd_init_kernel");
    const dim3 g1(ceil_div(n, EMD_THREADS), B), g2(ceil_div(m, EMD_THREADS), B);
    for (int j = 7; j >= -2; --j) {
        float level = -powf(4.0f, (float)j);  // emd_kernel.cu:51-54
        if (j == -2) level = 0.f;
        emd_rows1_kernel<1, 1><<<g1, EMD_THREADS, 0, st>>>(n, m, level, xyz1, xyz2, match, temp, nullptr);
        emd_rows2_kernel<1><<<g2, EMD_THREADS, 0, st>>>(n, m, level, xyz1, xyz2, temp);
        emd_rows1_kernel<3, 1><<<g1, EMD_THREADS, 0, st>>>(n, m, level, xyz1, xyz2, match, temp, nullptr);
    }
    B200PCI_LAUNCH_CHECK("emd sweep kernels");
    return B200PCI_OK;
}

extern "C" int b200pci_emd_matchcost(int B, int n, int m, const float *xyz1, const float *xyz2,
                                     const float *match, float *cost, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200PCI_CHECK_ARG(B >= 0 && n >= 0 && m >= 0, "emd_matchcost: negative size");
    if (B == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(cost, "emd_matchcost: null pointer");
    if (n == 0 || m == 0) {
        B200PCI_CUDA(cudaMemsetAsync(cost, 0, (size_t)B * sizeof(float), st));
        return B200PCI_OK;
    }
    B200PCI_CHECK_ARG(xyz1 && xyz2 && match, "emd_matchcost: null pointer");
    const size_t need = align_up((size_t)B * n * sizeof(float), 256);
    if (!workspace || workspace_bytes < need) {
        set_error("emd_matchcost: workspace of %zu bytes required, got %zu", need, workspace_bytes);
        return B200PCI_EWORKSPACE;
    }
    float *rowsum = reinterpret_cast<float *>(workspace);
    emd_cost_rows_kernel<<<dim3(ceil_div(n, EMD_THREADS), B), EMD_THREADS, 0, st>>>(n, m, xyz1, xyz2,
                                                                                  match, rowsum);
    emd_cost_reduce_kernel<float><<<B, 1024, 0, st>>>(n, rowsum, cost);
    B200PCI_LAUNCH_CHECK("emd_cost kernels");
    return B200PCI_OK;
}

static int g_emd_rs = 16;  // key 19 (developer): lanes per row of the forward-only sweeps (4, 8, 16, 32)
int b200pci_emd_debug_set(int key, double value) {
    if (key != 19) return B200PCI_EINVAL;
    g_emd_rs = (int)value;
    return B200PCI_OK;
}

template <int RS>
static void emd_cost_sweeps(int B, int n, int m, const float *xyz1, const float *xyz2, float *temp,
                            double *rowcost, cudaStream_t st) {
    using G = EmdGeom<RS>;
    const dim3 g1(ceil_div(n, G::rows), B), g2(ceil_div(m, G::rows), B);
    for (int j = 7; j >= -2; --j) {
        float level = -powf(4.0f, (float)j);  // emd_kernel.cu:51-54
        if (j == -2) level = 0.f;
        if (RS > 1) level *= 1.4426950408889634f;  // emd_exp<false>: exponent base 2
        emd_rows1_kernel<1, RS><<<g1, G::threads, 0, st>>>(n, m, level, xyz1, xyz2, nullptr, temp, nullptr);
        emd_rows2_kernel<RS><<<g2, G::threads, 0, st>>>(n, m, level, xyz1, xyz2, temp);
        emd_rows1_kernel<4, RS><<<g1, G::threads, 0, st>>>(n, m, level, xyz1, xyz2, nullptr, temp, rowcost);
    }
}

// Forward-only EMD (what the eval metric models/utils.py:223-235 needs): approxmatch + matchcost
// without ever storing `match` (1.07 GB per pair at 16384 x 16384, read-modify-written by every
// level of the reference, emd_kernel.cu:125-158).
extern "C" int b200pci_emd_cost(int B, int n, int m, const float *xyz1, const float *xyz2, float *cost,
                                void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200PCI_CHECK_ARG(B >= 0 && n >= 0 && m >= 0, "emd_cost: negative size");
    if (B == 0) return B200PCI_OK;
    B200PCI_CHECK_ARG(cost, "emd_cost: null pointer");
    if (n == 0 || m == 0) {
        B200PCI_CUDA(cudaMemsetAsync(cost, 0, (size_t)B * sizeof(float), st));
        return B200PCI_OK;
    }
    B200PCI_CHECK_ARG(xyz1 && xyz2, "emd_cost: null pointer");
    B200PCI_CHECK_ARG(B <= 65535, "emd_cost: batch too large");
    const size_t need = b200pci_emd_workspace_bytes(B, n, m);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 7)) {
        set_error("emd_cost: workspace of %zu bytes (8-B aligned) required, got %zu", need, workspace_bytes);
        return B200PCI_EWORKSPACE;
    }
    float *temp = reinterpret_cast<float *>(workspace);
    double *rowcost = reinterpret_cast<double *>(reinterpret_cast<char *>(workspace) + emd_temp_bytes(B, n, m));
    const float multiL = (n >= m) ? 1.f : (float)(m / n);
    const float multiR = (n >= m) ? (float)(n / m) : 1.f;
    B200PCI_CUDA(cudaMemsetAsync(rowcost, 0, (size_t)B * n * sizeof(double), st));
    const int mx = n > m ? n : m;
    emd_init_kernel<<<dim3(ceil_div(mx, 256), B), 256, 0, st>>>(n, m, multiL, multiR, temp);
    B200PCI_LAUNCH_CHECK("emd_init_kernel");
    switch (g_emd_rs) {
        case 4: emd_cost_sweeps<4>(B, n, m, xyz1, xyz2, temp, rowcost, st); break;
        case 32: emd_cost_sweeps<32>(B, n, m, xyz1, xyz2, temp, rowcost, st); break;
        case 8: emd_cost_sweeps<8>(B, n, m, xyz1, xyz2, temp, rowcost, st); break;
        default: emd_cost_sweeps<16>(B, n, m, xyz1, xyz2, temp, rowcost, st); break;
    }
    B200PCI_LAUNCH_CHECK("emd sweep kernels");
    emd_cost_reduce_kernel<double><<<B, 1024, 0, st>>>(n, rowcost, cost);
    B200PCI_LAUNCH_CHECK("emd_cost_reduce_kernel");
    return B200PCI_OK;
}

extern "C" int b200pci_emd_matchcost_grad(int B, int n, int m, const float *grad_cost,
                                          const float *xyz1, const float *xyz2, const float *match,
                                          float *grad1, float *grad2, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200PCI_CHECK_ARG(B >= 0 && n >= 0 && m >= 0, "emd_matchcost_grad: negative size");
    if (B == 0 || (n == 0 && m == 0)) return B200PCI_OK;
    B200PCI_CHECK_ARG(grad_cost && xyz1 && xyz2 && grad1 && grad2, "emd_matchcost_grad: null pointer");
    if (n == 0 || m == 0) {
        if (n) B200PCI_CUDA(cudaMemsetAsync(grad1, 0, (size_t)B * n * 3 * sizeof(float), st));
        if (m) B200PCI_CUDA(cudaMemsetAsync(grad2, 0, (size_t)B * m * 3 * sizeof(float), st));
        return B200PCI_OK;
    }
    B200PCI_CHECK_ARG(match, "emd_matchcost_grad: null pointer");
    emd_grad1_kernel<<<dim3(ceil_div(n, EMD_THREADS), B), EMD_THREADS, 0, st>>>(n, m, grad_cost, xyz1,
                                                                              xyz2, match, grad1);
    emd_grad2_kernel<<<dim3(ceil_div(m, 8), B), 256, 0, st>>>(n, m, grad_cost, xyz1, xyz2, match, grad2);
    B200PCI_LAUNCH_CHECK("emd grad kernels");
    return B200PCI_OK;
}
