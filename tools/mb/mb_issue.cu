// Micro-benchmark (developer tool): does an FFMA2 block the issue port for two cycles, i.e. can
// ALU-pipe (FMNMX) and LSU (broadcast LDS.128) instructions be issued "under" a stream of FFMA2?
// Prints cycles per loop iteration per SM sub-partition for several instruction mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mb_issue tools/mb/mb_issue.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fmaf_v(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmin_v(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// NF2 FFMA2 + NF scalar FFMA + NM FMNMX + NL broadcast LDS.128 per iteration
template <int NF2, int NF, int NM, int NL>
__global__ void __launch_bounds__(256) mix_kernel(int iters, float *sink, long long *cycles) {
    __shared__ float4 sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    __syncthreads();
    f32x2 acc2[NF2 > 0 ? NF2 : 1];
    float acc[NF > 0 ? NF : 1], mn[NM > 0 ? NM : 1];
    const float af = 1.0f + 1e-7f * threadIdx.x, bf = 1e-9f * blockIdx.x;
    const f32x2 a = pack2(af, af), b = pack2(bf, bf);
    for (int i = 0; i < NF2; ++i) acc2[i] = pack2((float)i, i + 0.5f);
    for (int i = 0; i < NF; ++i) acc[i] = (float)i;
    for (int i = 0; i < NM; ++i) mn[i] = 1e30f - i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float4 v[NL > 0 ? NL : 1];
#pragma unroll
        for (int i = 0; i < NL; ++i)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"((unsigned)__cvta_generic_to_shared(&sh[(it + i) & 63])));
#pragma unroll
        for (int i = 0; i < (NF2 > NM ? NF2 : NM); ++i) {
            if (i < NF2) {
                // the loaded values are the addends of the first 2*NL FFMA2 (keeps LDS.128 alive for free)
                f32x2 c = b;
                if (i < 2 * NL) c = (i & 1) ? pack2(v[i / 2].z, v[i / 2].w) : pack2(v[i / 2].x, v[i / 2].y);
                acc2[i] = fma2(acc2[i], a, c);
            }
            if (i < NM) mn[i] = fmin_v(mn[i], af + (float)(it & 7));
            if (i < NF) acc[i] = fmaf_v(acc[i], af, bf);
            if (NF > NF2 && i + 16 < NF) acc[i + 16] = fmaf_v(acc[i + 16], af, bf);
        }
        if (NF2 == 0 && NL > 0) {
#pragma unroll
            for (int i = 0; i < NL; ++i) mn[0] = fmin_v(fmin_v(mn[0], v[i].x + v[i].y), v[i].z + v[i].w);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < NF2; ++i) { float lo, hi; unpack2(acc2[i], lo, hi); s += lo + hi; }
    for (int i = 0; i < NF; ++i) s += acc[i];
    for (int i = 0; i < NM; ++i) s += mn[i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NF2, int NF, int NM, int NL>
void run(const char *name, int warps_per_smsp) {
    float *sink; long long *cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int ctas_per_sm = warps_per_smsp * 4 / 8;  // 256 threads = 8 warps
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix_kernel<NF2, NF, NM, NL><<<sms * ctas_per_sm, 256>>>(iters, sink, cyc);
    cudaEventRecord(e0);
    mix_kernel<NF2, NF, NM, NL><<<sms * ctas_per_sm, 256>>>(iters, sink, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double cyc_total = ms * 1e-3 * khz * 1e3;
    printf("%-34s w/smsp=%2d  events: %7.2f SMSP-cyc per warp-iteration | clock64: %7.2f cyc/iter/warp  (issue slots %d, fma-lane-cycles %d) %s\n",
           name, warps_per_smsp, cyc_total / iters / warps_per_smsp, (double)c / iters, NF2 + NF + NM + NL, 2 * NF2 + NF,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    for (int w : {2, 4, 8}) {
        if (w == 2) { run<16, 0, 0, 0>("16 FFMA2", 2); run<16, 0, 16, 0>("16 FFMA2 + 16 FMNMX", 2); run<0, 32, 16, 0>("32 FFMA + 16 FMNMX", 2); run<16,0,16,4>("16 FFMA2 + 16 FMNMX + 4 LDS.128", 2); run<16,0,8,0>("16 FFMA2 + 8 FMNMX", 2); run<0,32,0,0>("32 FFMA", 2); run<16,0,0,4>("16 FFMA2 + 4 LDS.128", 2);}
        if (w == 4) { run<16, 0, 0, 0>("16 FFMA2", 4); run<16, 0, 16, 0>("16 FFMA2 + 16 FMNMX", 4); run<0, 32, 16, 0>("32 FFMA + 16 FMNMX", 4); run<16,0,16,4>("16 FFMA2 + 16 FMNMX + 4 LDS.128", 4); run<16,0,8,0>("16 FFMA2 + 8 FMNMX", 4); run<0,32,0,0>("32 FFMA", 4); run<16,0,0,4>("16 FFMA2 + 4 LDS.128", 4);}
        if (w == 8) { run<16, 0, 0, 0>("16 FFMA2", 8); run<16, 0, 16, 0>("16 FFMA2 + 16 FMNMX", 8); run<0, 32, 16, 0>("32 FFMA + 16 FMNMX", 8); run<16,0,16,4>("16 FFMA2 + 16 FMNMX + 4 LDS.128", 8); run<16,0,8,0>("16 FFMA2 + 8 FMNMX", 8); run<0,32,0,0>("32 FFMA", 8); run<16,0,0,4>("16 FFMA2 + 4 LDS.128", 8); run<0,0,16,0>("16 FMNMX", 8); }
    }
    return 0;
}
