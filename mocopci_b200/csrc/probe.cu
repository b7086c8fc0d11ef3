// Roofline probe: measured FP32 FMA throughput (MEASURED_PEAKS.json has no FP32 figure).
// Each lane runs 16 independent FMA chains for `iters` rounds; scalar FFMA (packed=0) or the
// sm_100 packed FFMA2 (packed=1). bench.py times the launch with CUDA events.
#include "common.cuh"

namespace b200pci {

constexpr int PROBE_CHAINS = 16;
constexpr int PROBE_THREADS = 256;
constexpr int PROBE_CTAS_PER_SM = 8;

__global__ void __launch_bounds__(PROBE_THREADS) probe_ffma_kernel(int iters, float *sink) {
    float acc[PROBE_CHAINS];
    const float a = 1.0f + 1e-7f * threadIdx.x, b = 1e-9f * blockIdx.x;
#pragma unroll
    for (int i = 0; i < PROBE_CHAINS; ++i) acc[i] = (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < PROBE_CHAINS; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PROBE_CHAINS; ++i) s += acc[i];
    if (s == 123.456f) sink[0] = s;  // keep the chains alive
}

__global__ void __launch_bounds__(PROBE_THREADS) probe_ffma2_kernel(int iters, float *sink) {
    f32x2 acc[PROBE_CHAINS];
    const float af = 1.0f + 1e-7f * threadIdx.x, bf = 1e-9f * blockIdx.x;
    const f32x2 a = pack2(af, af), b = pack2(bf, bf);
#pragma unroll
    for (int i = 0; i < PROBE_CHAINS; ++i) acc[i] = pack2((float)i, (float)i + 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < PROBE_CHAINS; ++i) acc[i] = fma2(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PROBE_CHAINS; ++i) {
        float lo, hi;
        unpack2(acc[i], lo, hi);
        s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
}

}  // namespace b200pci

using namespace b200pci;

extern "C" int b200pci_probe_fp32(int packed, int iters, float *sink, double *flops, void *stream) {
    B200PCI_CHECK_ARG(iters > 0 && sink, "probe_fp32: bad arguments");
    const int ctas = sm_count() * PROBE_CTAS_PER_SM;
    if (packed)
        probe_ffma2_kernel<<<ctas, PROBE_THREADS, 0, (cudaStream_t)stream>>>(iters, sink);
    else
        probe_ffma_kernel<<<ctas, PROBE_THREADS, 0, (cudaStream_t)stream>>>(iters, sink);
    B200PCI_LAUNCH_CHECK("probe_fp32");
    if (flops)
        *flops = 2.0 * (packed ? 2.0 : 1.0) * PROBE_CHAINS * (double)iters * PROBE_THREADS * ctas;
    return B200PCI_OK;
}
