// Micro-benchmark (developer tool): FFMA2 throughput vs operand pattern (register-file bandwidth).
//   A: acc = fma2(acc, a, b)           a, b shared by all chains (operand reuse cache hits)
//   B: acc_i = fma2(x_i, q_j, acc_i)   three distinct register operands (64 + 32 + 64 bit)
//   C: like B but 4 consecutive FFMA2 share x (the scan loop's ideal order: same refs, 4 queries)
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, float *sink, const float *in) {
    f32x2 acc[16], x[16];
    float qv[4];
    for (int i = 0; i < 16; ++i) { acc[i] = pack2((float)i, i + 0.5f); x[i] = pack2(in[i] + threadIdx.x, in[i + 16]); }
    for (int i = 0; i < 4; ++i) qv[i] = in[32 + i] * (1.f + threadIdx.x * 1e-7f);
    const f32x2 a = pack2(qv[0], qv[0]), b = pack2(qv[1], qv[1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) acc[i] = fma2(acc[i], a, b);
            if (MODE == 1) acc[i] = fma2(x[(i * 5) & 15], pack2(qv[i & 3], qv[i & 3]), acc[i]);
            if (MODE == 2) acc[i] = fma2(x[i >> 2], pack2(qv[i & 3], qv[i & 3]), acc[i]);
        }
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) { float lo, hi; unpack2(acc[i], lo, hi); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}
template <int MODE> void run(const char *name) {
    float *sink, *in; cudaMalloc(&sink, 4); cudaMalloc(&in, 256); cudaMemset(in, 0, 256);
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int iters = 20000, wps = 4;  // 4 warps per sub-partition
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 2, 256>>>(iters, sink, in);
    cudaEventRecord(e0); k<MODE><<<sms * 2, 256>>>(iters, sink, in); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-60s %.2f cycles per FFMA2 per sub-partition (%s)\n", name, ms * 1e-3 * khz * 1e3 / iters / wps / 16, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("A acc=fma2(acc,a,b), shared a,b");
    run<1>("B acc_i=fma2(x_k,q_j,acc_i), distinct operands");
    run<2>("C acc_i=fma2(x_(i/4),q_j,acc_i), x shared by 4 consecutive");
    return 0;
}
