// Host-side int32 -> int64 widening rate with T threads (is "copy int32 over PCIe, widen on the CPU"
// faster than copying int64?):  gcc -O3 -march=native -pthread mb_widen.c -o mb_widen && ./mb_widen
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { const int32_t *src; int64_t *dst; size_t n; } job_t;
static void *work(void *p) {
    job_t *j = (job_t *)p;
    for (size_t i = 0; i < j->n; ++i) j->dst[i] = j->src[i];
    return NULL;
}
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(void) {
    const size_t n = (size_t)64 * 16384 * 16;
    int32_t *src = aligned_alloc(4096, n * 4);
    int64_t *dst = aligned_alloc(4096, n * 8);
    for (size_t i = 0; i < n; ++i) src[i] = (int32_t)(i & 16383);
    memset(dst, 0, n * 8);
    for (int T = 1; T <= 32; T *= 2) {
        pthread_t th[32]; job_t jobs[32];
        double best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            const double t0 = now();
            for (int t = 0; t < T; ++t) {
                const size_t a = n * t / T, b = n * (t + 1) / T;
                jobs[t].src = src + a; jobs[t].dst = dst + a; jobs[t].n = b - a;
                pthread_create(&th[t], NULL, work, &jobs[t]);
            }
            for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
            const double dt = now() - t0;
            if (dt < best) best = dt;
        }
        printf("threads %2d: %.3f ms for %zu MB out (%.1f GB/s written)\n", T, best * 1e3, n * 8 >> 20, n * 8 / best / 1e9);
    }
    return 0;
}
