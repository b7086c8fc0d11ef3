// Spatial ordering for the big neighbourhood searches (two-pass tensor-core path, clouds of up to
// 16384 points): points are sorted along a 30-bit Morton curve so that 128 consecutive points form
// a compact blob with a small bounding box. The scan (nbr_scan_tc.cuh) then skips every ref tile
// whose box is farther from the CTA's query box than the largest admission bound of its queries --
// such a tile cannot hold a candidate of any of them -- instead of pushing it through the filter.
// RESULTS DO NOT CHANGE: candidate keys carry the ORIGINAL ref index (so ties still go to the
// lowest original index) and result rows are written back to the original query positions; the
// ordering only decides which (query tile, ref tile) pairs are looked at.
//
//   nbr_sort_kernel   one CTA of 1024 threads per cloud: bounding box (block reduce), Morton keys
//                     (code << 32 | index) in shared memory, bitonic sort, then it writes the
//                     permutation (sorted position -> original index), the sorted AoS copy of the
//                     cloud and the bounding box of every 128-point block.
#pragma once
#include "common.cuh"

namespace b200pci {

constexpr int SORT_MAX_POINTS = 16384;  // keys of one cloud live in shared memory (8 B each)
constexpr int SORT_THREADS = 1024;
constexpr int SORT_BLOCK = 128;         // points per bounding box (= a ref tile of the scan)

struct SortCloud {
    const float *src;          // (b, i, c) at src[b*sb + i*sp + c*sc]
    long long sb, sp, sc;
    int n;
    int *perm;                 // [B][n] sorted position -> original index
    float *sorted;             // [B][n][3] contiguous
    float *boxes;              // [B][ceil(n / 128)][8]: lo.xyz, hi.xyz, max |p|^2, unused
};

__device__ __forceinline__ uint32_t morton_expand10(uint32_t v) {  // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// grid (clouds, B); dynamic shared memory: P * 8 bytes, P = n rounded up to a power of two
static __global__ void __launch_bounds__(SORT_THREADS, 1) nbr_sort_kernel(SortCloud c0, SortCloud c1) {
    extern __shared__ __align__(16) unsigned long long skey[];
    __shared__ float red[6][32];
    __shared__ float bb[6];
    const SortCloud c = blockIdx.x == 0 ? c0 : c1;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = c.n;
    const float *src = c.src + b * c.sb;
    int P = 1;
    while (P < n) P <<= 1;
    // ---- bounding box of the cloud ----
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = tid; i < n; i += SORT_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = src[i * c.sp + a * c.sc];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min_f(lo[a]);
        hi[a] = warp_max_f(hi[a]);
        if (lane == 0) {
            red[a][warp] = lo[a];
            red[3 + a][warp] = hi[a];
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float l = warp_min_f(red[a][lane]), h = warp_max_f(red[3 + a][lane]);
            if (lane == 0) {
                bb[a] = l;
                bb[3 + a] = h;
            }
        }
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float ext = bb[3 + a] - bb[a];
        scale[a] = (ext > 0.f && ext < inf) ? 1023.0f / ext : 0.f;
    }
    // ---- keys ----
    for (int i = tid; i < P; i += SORT_THREADS) {
        unsigned long long key = ~0ull;  // padding sorts last
        if (i < n) {
            uint32_t code = 0u;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = src[i * c.sp + a * c.sc];
                float t = (v - bb[a]) * scale[a];
                t = fminf(fmaxf(t, 0.f), 1023.f);  // (NaN -> 0)
                code |= morton_expand10((uint32_t)t) << a;
            }
            key = ((unsigned long long)code << 32) | (uint32_t)i;
        }
        skey[i] = key;
    }
    __syncthreads();
    // ---- bitonic sort (ascending) ----
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P / 2; t += SORT_THREADS) {
                // the t-th pair of this stage: insert a zero bit at position log2(j)
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const unsigned long long a = skey[i], bq = skey[l];
                const bool up = (i & k) == 0;
                if ((a > bq) == up) {
                    skey[i] = bq;
                    skey[l] = a;
                }
            }
            __syncthreads();
        }
    }
    // ---- outputs: permutation, sorted copy, block boxes ----
    int *perm = c.perm + (size_t)b * n;
    float *sorted = c.sorted + (size_t)b * n * 3;
    const int nblocks = (n + SORT_BLOCK - 1) / SORT_BLOCK;
    float *boxes = c.boxes + (size_t)b * nblocks * 8;
    for (int blk = warp; blk < nblocks; blk += SORT_THREADS / 32) {
        float l3[3] = {inf, inf, inf}, h3[3] = {-inf, -inf, -inf}, nmax = 0.f;
#pragma unroll
        for (int u = 0; u < SORT_BLOCK / 32; ++u) {
            const int pos = blk * SORT_BLOCK + u * 32 + lane;
            if (pos < n) {
                const int orig = (int)(uint32_t)skey[pos];
                perm[pos] = orig;
                float v[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    v[a] = src[orig * c.sp + a * c.sc];
                    sorted[(size_t)pos * 3 + a] = v[a];
                    l3[a] = fminf(l3[a], v[a]);
                    h3[a] = fmaxf(h3[a], v[a]);
                }
                nmax = fmaxf(nmax, fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0])));
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            l3[a] = warp_min_f(l3[a]);
            h3[a] = warp_max_f(h3[a]);
        }
        nmax = warp_max_f(nmax);
        if (lane == 0) {
            float *o = boxes + (size_t)blk * 8;
            o[0] = l3[0];
            o[1] = l3[1];
            o[2] = l3[2];
            o[3] = h3[0];
            o[4] = h3[1];
            o[5] = h3[2];
            o[6] = nmax;
            o[7] = 0.f;
        }
    }
}

// squared distance between two axis-aligned boxes (0 if they overlap); lo/hi as in SortCloud::boxes
__device__ __forceinline__ float box_gap2(const float *a, const float *b) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float g = fmaxf(fmaxf(a[k] - b[3 + k], b[k] - a[3 + k]), 0.f);
        s = fmaf(g, g, s);
    }
    return s;
}

}  // namespace b200pci
