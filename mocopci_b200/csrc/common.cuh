// Shared helpers for the b200pci kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200pci.h"

namespace b200pci {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define B200PCI_CHECK_ARG(cond, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            ::b200pci::set_error(__VA_ARGS__); \
            return B200PCI_EINVAL;            \
        }                                     \
    } while (0)

#define B200PCI_CUDA(call)                                              \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::b200pci::cuda_fail(e__, #call); \
    } while (0)

#define B200PCI_LAUNCH_CHECK(name)                                        \
    do {                                                                  \
        cudaError_t e__ = cudaGetLastError();                             \
        if (e__ != cudaSuccess) return ::b200pci::cuda_fail(e__, name);    \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
int sm_count();

// ---- division of a 32-bit index by a run-time constant: one IMAD.HI + shift instead of the ~20
// (32-bit) / ~60 (64-bit) instructions of a hardware-less integer divide. Built on the host.
struct FastDiv {
    uint32_t d, mul, shift;
};
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    f.shift = 0;
    while ((1ull << f.shift) < d) ++f.shift;
    // q = (umulhi(n, mul) + n) >> shift  for every n < 2^31  (Granlund-Montgomery / Hacker's Delight 10-9)
    f.mul = (uint32_t)((((1ull << f.shift) - d) << 32) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv f) {
    return (uint32_t)(((unsigned long long)__umulhi(n, f.mul) + n) >> f.shift);
}

// ---- packed FP32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2, IEEE round-to-nearest per lane) --
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- order-preserving float <-> uint32 (handles the slightly negative expanded distances) -----
__device__ __forceinline__ uint32_t f2sortable(float f) {
    uint32_t b = __float_as_uint(f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float sortable2f(uint32_t s) {
    uint32_t b = s ^ ((s & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long make_key(float d, uint32_t idx) {
    return ((unsigned long long)f2sortable(d) << 32) | idx;
}
#define B200PCI_KEY_INF 0xFF80000000000000ull /* (+inf, idx 0): empty slot */

// ---- mbarrier + 1-D TMA bulk copy -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(bytes)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// the same for single-thread producer / issuer loops: the thread is suspended (up to the hint, ns)
// instead of spinning on issue slots the other warps of its scheduler could use
__device__ __forceinline__ void mbar_wait_suspend(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (ok == 0);
}
// global -> shared bulk copy (TMA, SASS UBLKCP), completion counted on `bar` in bytes.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 256-bit read-only global load (sm_100: LDG.E.256), p 32-byte aligned
__device__ __forceinline__ void ldg256(const float *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

__device__ __forceinline__ int warp_max_i(int v) {
    return __reduce_max_sync(0xffffffffu, v);
}

}  // namespace b200pci
