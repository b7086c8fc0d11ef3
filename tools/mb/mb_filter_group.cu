// Micro-benchmark (developer tool): the scan's per-group instruction pattern in isolation --
// 4 broadcast LDS.128 (next group), 24 FFMA2 (8 chains of 3), 4 x (FMNMX, FMNMX3, FSETP, predicated
// mask update) -- to see what the pattern itself costs per group on an SM sub-partition.
//   variants: full / no compare (FFMA2 + LDS only) / no LDS (registers only) /
//             step-min (running minimum over the 8 groups of a step, one compare per step)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>  // 0 full, 1 no compare, 2 no LDS, 3 step-min
__global__ void __launch_bounds__(32, 16) k(int iters, const float *in, unsigned *out) {
    __shared__ float4 sm[4 * 64];
    for (int i = threadIdx.x; i < 256; i += 32) sm[i] = make_float4(in[i & 63], in[(i + 1) & 63], in[(i + 2) & 63], in[(i + 3) & 63]);
    __syncwarp();
    float fa[4], fb[4], fc[4], thr[4];
    for (int j = 0; j < 4; ++j) { fa[j] = in[j] + threadIdx.x; fb[j] = in[4 + j]; fc[j] = in[8 + j]; thr[j] = in[12 + j] - 1e30f; }
    unsigned m8[4] = {0, 0, 0, 0};
    float4 X = sm[0], Y = sm[64], Z = sm[128], W = sm[192];
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float smin[4] = {1e30f, 1e30f, 1e30f, 1e30f};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 cX = X, cY = Y, cZ = Z, cW = W;
            if (MODE != 2) {
                const int g = (it * 8 + u + 1) & 63;
                X = sm[g]; Y = sm[64 + g]; Z = sm[128 + g]; W = sm[192 + g];
            } else {
                X.x += 1.f;  // keep a dependency so the loop is not hoisted
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f32x2 a = pack2(fa[j], fa[j]), b = pack2(fb[j], fb[j]), c = pack2(fc[j], fc[j]);
                f32x2 t0 = fma2(pack2(cX.x, cX.y), a, pack2(cW.x, cW.y));
                f32x2 t1 = fma2(pack2(cX.z, cX.w), a, pack2(cW.z, cW.w));
                t0 = fma2(pack2(cY.x, cY.y), b, t0);
                t1 = fma2(pack2(cY.z, cY.w), b, t1);
                t0 = fma2(pack2(cZ.x, cZ.y), c, t0);
                t1 = fma2(pack2(cZ.z, cZ.w), c, t1);
                float d0, d1, d2, d3;
                unpack2(t0, d0, d1);
                unpack2(t1, d2, d3);
                if (MODE == 1) {
                    acc += d0 + d2;  // 2 FADD instead of min/compare/mask
                } else if (MODE == 3) {
                    smin[j] = fminf(fminf(smin[j], d0), d1);
                    smin[j] = fminf(fminf(smin[j], d2), d3);
                } else {
                    if (fminf(fminf(d0, d1), fminf(d2, d3)) < thr[j]) m8[j] |= (0x80u >> u);
                }
            }
        }
        if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (smin[j] < thr[j]) m8[j] |= 1u;
        }
        if (m8[0] | m8[1] | m8[2] | m8[3]) { out[0] = m8[0]; m8[0] = 0; }
    }
    if (acc == 123.f || m8[1] == 77u) out[1] = 1;
}
template <int MODE> void run(const char *name, int ctas_per_sm) {
    float *in; unsigned *out; cudaMalloc(&in, 1024); cudaMalloc(&out, 16);
    float h[256]; for (int i = 0; i < 256; ++i) h[i] = 0.5f + 0.01f * i; cudaMemcpy(in, h, 1024, cudaMemcpyHostToDevice);
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * ctas_per_sm, 32>>>(iters, in, out);
    cudaEventRecord(e0); k<MODE><<<sms * ctas_per_sm, 32>>>(iters, in, out); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps_per_smsp = ctas_per_sm / 4.0;
    printf("%-34s warps/SMSP=%.1f  %.1f cycles per group per SMSP-warp  (%s)\n", name, warps_per_smsp,
           ms * 1e-3 * khz * 1e3 / (iters * 8.0) / warps_per_smsp, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    for (int c : {4, 8, 14, 16}) {
        if (c == 4) { run<0>("full pattern", 4); run<1>("FFMA2 + LDS (no compare)", 4); run<3>("step-min", 4); }
        if (c == 8) { run<0>("full pattern", 8); run<1>("FFMA2 + LDS (no compare)", 8); run<3>("step-min", 8); }
        if (c == 14) { run<0>("full pattern", 14); run<1>("FFMA2 + LDS (no compare)", 14); run<3>("step-min", 14); }
        if (c == 16) { run<0>("full pattern", 16); run<1>("FFMA2 + LDS (no compare)", 16); run<3>("step-min", 16); }
    }
    return 0;
}
