// Developer tool: one 128 x 128 x 16 tcgen05 TF32 tile of the tensor-core filter
//   D[m][n] = sum_k A[m][k] * B[n][k]   ~   w'_n - 2 q_m . r_n   (split-TF32 operands)
// from K-major, no-swizzle shared-memory operands laid out [k/4][row][4 floats]; prints the error
// against a double-precision evaluation in units of (|q|^2 + |r|^2), for two descriptor
// conventions (which of LBO / SBO is the K-direction stride), and an MMA issue-rate figure.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (long long i = 0; i < 200000000ll; ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);  // F32 acc, TF32 x TF32, K-major, N=128, M=128
__device__ __forceinline__ uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

// A, B: [128][16] row-major in global; D out [128][128]; iters > 1: repeat the MMAs (timing)
__global__ void __launch_bounds__(128) tile_kernel(const float *A, const float *B, float *D, uint32_t lbo, uint32_t sbo,
                                                   int iters, long long *cycles, int *status, int ncols = 128, int nacc = 1) {
    __shared__ __align__(128) float sA[4 * 128 * 4];
    __shared__ __align__(128) float sB[4 * 128 * 4];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int c = 0; c < 4; ++c) {
        reinterpret_cast<float4 *>(sA)[c * 128 + tid] = reinterpret_cast<const float4 *>(A)[tid * 4 + c];
        reinterpret_cast<float4 *>(sB)[c * 128 + tid] = reinterpret_cast<const float4 *>(B)[tid * 4 + c];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_slot;
    long long t0 = 0;
    if (tid == 0) {
        t0 = clock64();
        const uint32_t id = idesc_n(ncols);
        for (int it = 0; it < iters; ++it) {
            // K elements 0..7 = chunks 0,1; 8..15 = chunks 2,3 (chunk stride 128 rows x 16 B)
            const uint32_t d = taddr + (uint32_t)(it % nacc) * ncols;
            mma_tf32(d, make_desc(smem_u32(sA), lbo, sbo), make_desc(smem_u32(sB), lbo, sbo), id, 0u);
            mma_tf32(d, make_desc(smem_u32(sA) + 4096, lbo, sbo), make_desc(smem_u32(sB) + 4096, lbo, sbo), id, 1u);
        }
        mma_commit(&bar);
    }
    __syncwarp();
    const bool ok = mbar_wait_bounded(&bar, 0);
    if (tid == 0) {
        *cycles = clock64() - t0;
        *status = ok ? 0 : 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        for (int step = 0; step < 4; ++step) {
            uint32_t v[32];
            tmem_ld32(taddr + ((uint32_t)(warp * 32) << 16) + step * 32, v);
            for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + lane) * 128 + step * 32 + i] = __uint_as_float(v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(128u) : "memory");
}

static float tf32_trunc(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u &= 0xFFFFE000u;
    memcpy(&f, &u, 4);
    return f;
}
static void split2(float v, float &h, float &l) {
    h = tf32_trunc(v);
    l = tf32_trunc(v - h);
}

int main() {
    std::vector<float> A(128 * 16), B(128 * 16), D(128 * 128);
    std::vector<double> qx(128), qy(128), qz(128), rx(128), ry(128), rz(128), w(128);
    srand(7);
    auto rnd = []() { return (float)((rand() / (double)RAND_MAX) * 100.0 - 50.0); };
    for (int i = 0; i < 128; ++i) {
        const float q[3] = {rnd(), rnd(), rnd()}, r[3] = {rnd(), rnd(), rnd()};
        qx[i] = q[0]; qy[i] = q[1]; qz[i] = q[2];
        rx[i] = r[0]; ry[i] = r[1]; rz[i] = r[2];
        const float wn = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        w[i] = wn;
        float *a = &A[i * 16], *b = &B[i * 16];
        for (int c = 0; c < 3; ++c) {
            float qh, ql, rh, rl;
            split2(-2.f * q[c], qh, ql);
            split2(r[c], rh, rl);
            a[c] = qh; b[c] = rh;          // qh * rh
            a[4 + c] = qh; b[4 + c] = rl;  // qh * rl
            a[8 + c] = ql; b[8 + c] = rh;  // ql * rh
            a[12 + c] = ql; b[12 + c] = rl;
        }
        float wh, wl, wll;
        wh = tf32_trunc(wn);
        wl = tf32_trunc(wn - wh);
        wll = tf32_trunc((wn - wh) - wl);
        a[3] = 1.f; b[3] = wh;
        a[7] = 1.f; b[7] = wl;
        a[11] = 1.f; b[11] = wll;
        a[15] = 0.f; b[15] = 0.f;
    }
    float *dA, *dB, *dD;
    long long *dc;
    int *ds;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&dD, D.size() * 4);
    cudaMalloc(&dc, 8);
    cudaMalloc(&ds, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const uint32_t conv[1][2] = {{2048u, 128u}};  // {LBO, SBO}; the swapped convention faults
    for (int v = 0; v < 1; ++v) {
        cudaMemset(dD, 0, D.size() * 4);
        tile_kernel<<<1, 128>>>(dA, dB, dD, conv[v][0], conv[v][1], 1, dc, ds);
        cudaError_t e = cudaDeviceSynchronize();
        int st = -1;
        cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0.0, worst_abs = 0.0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                const double ref = w[n] - 2.0 * (qx[m] * rx[n] + qy[m] * ry[n] + qz[m] * rz[n]);
                const double P = qx[m] * qx[m] + qy[m] * qy[m] + qz[m] * qz[m] + w[n];
                const double err = fabs((double)D[m * 128 + n] - ref);
                if (err / P > worst) worst = err / P;
                if (err > worst_abs) worst_abs = err;
            }
        printf("convention LBO=%u SBO=%u: cuda=%s status=%d  max |err| = %.3g, max |err|/(|q|^2+|r|^2) = %.3g = 2^%.1f  D[0][0]=%g D[5][77]=%g\n",
               conv[v][0], conv[v][1], cudaGetErrorString(e), st, worst_abs, worst, worst > 0 ? log2(worst) : -99.0, D[0], D[5 * 128 + 77]);
        if (e != cudaSuccess) return 1;
    }
    // issue rate: 2 MMAs (128x128x8 each) per iteration
    const int cfg[4][2] = {{128, 1}, {64, 1}, {64, 2}, {32, 4}};
    for (auto &c2 : cfg)
        for (int iters : {64, 1024}) {
            tile_kernel<<<1, 128>>>(dA, dB, dD, 2048u, 128u, iters, dc, ds, c2[0], c2[1]);
            cudaDeviceSynchronize();
            long long c = 0;
            cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            printf("N=%d, %d accumulator(s), iters %d: %lld cycles, %.1f cycles per 128xNx16 (two MMAs)\n", c2[0], c2[1], iters, c,
                   (double)c / iters);
        }
    return 0;
}
