/* TEST INFRASTRUCTURE ONLY -- see oracle.h. Plain C restatement of the reference algorithms;
 * build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -lm  (oracle/Makefile).
 * Paths below are relative to /root/reference. */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------- */
/* Distance forms                                                                            */
/* ---------------------------------------------------------------------------------------- */

/* models/pointconv_util.py:85-87 as executed by CPU torch (oneMKL sgemm, K=3) -- rounding
 * sequence established in SURVEY section 8a row K1 and re-checked by tests/golden/make_golden.py:
 *   dot = fma(z,Z, fma(y,Y, x*X));  s = (x*x + y*y) + z*z;  D = ((-2*dot) + s_src) + s_dst     */
static inline float sqnorm3(const float *p) { return (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]; }
/* The same sum as CUDA torch's reduction over a last dimension of 3 evaluates it: (a + c) + b
 * (measured on B200 with torch 2.11, tools/gpu_probe.py; everything else in square_distance --
 * the K = 3 matmul, the -2 scaling, the two in-place adds -- gives the same bits on CPU and CUDA). */
static inline float sqnorm3_cuda(const float *p) { return (p[0] * p[0] + p[2] * p[2]) + p[1] * p[1]; }
static inline float dist_expanded(const float *q, float sq, const float *r, float sr) {
    float dot = fmaf(q[2], r[2], fmaf(q[1], r[1], q[0] * r[0]));
    float t = -2.0f * dot + sq; /* -2*dot is exact, one rounding */
    return t + sr;
}

/* (a-b)^2 sum as nvcc -O2 contracts it in every reference kernel (SASS of the reference objects:
 * FMUL dy*dy; FFMA dx*dx+.; FFMA dz*dz+.) -- interpolate_gpu.cu:37, ball_query_gpu.cu:33,
 * sampling_gpu.cu:130, emd_kernel.cu:82. */
static inline float dist_direct(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* pytorch3d 0.7.5 (environment.yaml:90; NOT vendored -- restated from its published source,
 * pytorch3d/csrc/knn/knn.cu: `for d in 0..D-1: diff = p1[d] - p2[d]; dist += diff * diff`, compiled by
 * nvcc with -fmad=true): x first, then y, then z, each step one FMA. PARITY UNPINNED. */
static inline float dist_direct_xyz(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* models/pointT_layer2.py:20 torch.sum((src[:, :, None] - dst[:, None]) ** 2, dim=-1): elementwise
 * square, then a 3-element sum -- every operation rounded, no FMA. */
static inline float dist_sqdiff(float dx, float dy, float dz) {
    return (dx * dx + dy * dy) + dz * dz;
}
static inline float dist_sqdiff_cuda(float dx, float dy, float dz) { /* CUDA torch's sum order */
    return (dx * dx + dz * dz) + dy * dy;
}

/* form: 0 expanded, 1 pointnet2 (dist_direct), 2 pytorch3d (dist_direct_xyz), 3 pointT (dist_sqdiff),
 * 4 pointT as CUDA torch sums it, 5 expanded as CUDA torch sums the norms */
static inline float dist_form(int form, float dx, float dy, float dz) {
    return form == 2 ? dist_direct_xyz(dx, dy, dz) : form == 3 ? dist_sqdiff(dx, dy, dz)
           : form == 4 ? dist_sqdiff_cuda(dx, dy, dz) : dist_direct(dx, dy, dz);
}

void orc_square_distance(int B, int S, int N, const float *q, const float *r, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < S; ++i) {
            const float *qi = q + ((size_t)b * S + i) * 3;
            float sq = sqnorm3(qi);
            float *o = out + ((size_t)b * S + i) * N;
            for (int j = 0; j < N; ++j) {
                const float *rj = r + ((size_t)b * N + j) * 3;
                o[j] = dist_expanded(qi, sq, rj, sqnorm3(rj));
            }
        }
}

/* ---------------------------------------------------------------------------------------- */
/* k-NN selection: keep the k smallest by (distance, index); refs are visited in ascending    */
/* index order, so "strictly smaller than the current k-th" == lowest index wins ties.        */
/* ---------------------------------------------------------------------------------------- */
static inline void topk_insert(float *bd, int64_t *bi, int k, float d, int64_t j) {
    if (!(d < bd[k - 1])) return;
    int p = k - 1;
    while (p > 0 && bd[p - 1] > d) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        --p;
    }
    bd[p] = d;
    bi[p] = j;
}

static int knn_generic(int B, int S, int N, int k, const float *q, const float *r, int64_t *idx,
                       float *dist, int form) {
    const int expanded = form == 0 || form == 5;
    if (k <= 0) return -1;
    float *rn = NULL;
    if (expanded) {
        rn = (float *)malloc(sizeof(float) * (size_t)B * (N > 0 ? N : 1));
        for (size_t t = 0; t < (size_t)B * N; ++t) rn[t] = form == 5 ? sqnorm3_cuda(r + t * 3) : sqnorm3(r + t * 3);
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < S; ++i) {
            float bd[64];
            int64_t bi[64];
            float *pd = (k <= 64) ? bd : (float *)malloc(sizeof(float) * k);
            int64_t *pi = (k <= 64) ? bi : (int64_t *)malloc(sizeof(int64_t) * k);
            for (int t = 0; t < k; ++t) {
                pd[t] = INFINITY;
                pi[t] = 0;
            }
            const float *qi = q + ((size_t)b * S + i) * 3;
            float sq = form == 5 ? sqnorm3_cuda(qi) : sqnorm3(qi);
            for (int j = 0; j < N; ++j) {
                const float *rj = r + ((size_t)b * N + j) * 3;
                float d = expanded ? dist_expanded(qi, sq, rj, rn[(size_t)b * N + j])
                                   : dist_form(form, qi[0] - rj[0], qi[1] - rj[1], qi[2] - rj[2]);
                topk_insert(pd, pi, k, d, j);
            }
            for (int t = 0; t < k; ++t) {
                idx[((size_t)b * S + i) * k + t] = pi[t];
                if (dist) dist[((size_t)b * S + i) * k + t] = pd[t];
            }
            if (k > 64) {
                free(pd);
                free(pi);
            }
        }
    free(rn);
    return 0;
}

int orc_knn_expanded(int B, int S, int N, int k, const float *q, const float *r, int64_t *idx,
                     float *dist) {
    if (k > N) return -1; /* torch.topk raises for k > N */
    return knn_generic(B, S, N, k, q, r, idx, dist, 0);
}

int orc_knn_form(int form, int B, int S, int N, int k, const float *q, const float *r, int64_t *idx,
                 float *dist) {
    if (form < 0 || form > 5) return -1;
    if ((form == 0 || form == 5) && k > N) return -1;
    return knn_generic(B, S, N, k, q, r, idx, dist, form);
}

/* models/pointconv_util.py:111-127 cosine_distance + :142-153 knn_point_cosine.
 * q [B,S,C] = new_xyz, r [B,N,C] = xyz. Rows are normalised in float32 exactly as the reference
 * writes it (x / sqrt(sum(x^2) + 1e-8), the sum in double then rounded: torch's own reduction order
 * is unspecified), the dot product is accumulated in double and rounded once, dist = 1 - dot.
 * The reference's bmm (oneMKL / cuBLAS sgemm) sums in an unspecified order, so comparisons against
 * it use a tolerance (~C ulp of 1). Sorted by (distance, index). Returns -1 if k > N. */
int orc_knn_cosine(int B, int S, int N, int C, int k, const float *q, const float *r, int64_t *idx,
                   float *dist) {
    if (k <= 0 || k > N || k > 64) return -1;
    float *qn = (float *)malloc(sizeof(float) * (size_t)B * S * C);
    float *rn = (float *)malloc(sizeof(float) * (size_t)B * N * C);
    for (int pass = 0; pass < 2; ++pass) {
        const float *src = pass ? r : q;
        float *dst = pass ? rn : qn;
        const size_t rows = (size_t)B * (pass ? N : S);
#pragma omp parallel for schedule(static)
        for (size_t t = 0; t < rows; ++t) {
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += (double)(src[t * C + c] * src[t * C + c]);
            const float nrm = sqrtf((float)s + 1e-8f);
            for (int c = 0; c < C; ++c) dst[t * C + c] = src[t * C + c] / nrm;
        }
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < S; ++i) {
            float bd[64];
            int64_t bi[64];
            for (int t = 0; t < k; ++t) {
                bd[t] = INFINITY;
                bi[t] = 0;
            }
            const float *qi = qn + ((size_t)b * S + i) * C;
            for (int j = 0; j < N; ++j) {
                const float *rj = rn + ((size_t)b * N + j) * C;
                double dot = 0.0;
                for (int c = 0; c < C; ++c) dot += (double)qi[c] * (double)rj[c];
                topk_insert(bd, bi, k, 1.0f - (float)dot, j);
            }
            for (int t = 0; t < k; ++t) {
                idx[((size_t)b * S + i) * k + t] = bi[t];
                if (dist) dist[((size_t)b * S + i) * k + t] = bd[t];
            }
        }
    free(qn);
    free(rn);
    return 0;
}

/* pointnet2/pointnet2_modules.py:139-144 on top of pointnet2_utils.py:97 (T3):
 * dist = sqrt(dist2); r = 1/(dist + 1e-8); weight = r / sum(r) with the sum as (r0 + r2) + r1,
 * the order of CUDA torch's reduction (the reference's three_nn only exists on CUDA). */
void orc_three_nn_weights(long long rows, float eps, const float *dist2, float *dist, float *weight) {
    for (long long i = 0; i < rows; ++i) {
        float s[3], r[3];
        for (int j = 0; j < 3; ++j) {
            s[j] = sqrtf(dist2[i * 3 + j]);
            r[j] = 1.0f / (s[j] + eps);
        }
        const float norm = (r[0] + r[2]) + r[1]; /* CUDA torch's 3-element sum; three_nn is CUDA-only */
        for (int j = 0; j < 3; ++j) {
            dist[i * 3 + j] = s[j];
            weight[i * 3 + j] = r[j] / norm;
        }
    }
}

int orc_knn_direct(int B, int S, int N, int k, const float *q, const float *r, int64_t *idx,
                   float *dist) {
    return knn_generic(B, S, N, k, q, r, idx, dist, 1);
}

/* ---------------------------------------------------------------------------------------- */
/* FPS -- sampling_gpu.cu:93-209                                                             */
/* ---------------------------------------------------------------------------------------- */
static int ref_opt_n_threads(int work_size) { /* cuda_utils.h:10-14 */
    int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

void orc_fps(int B, int N, int M, const float *xyz, float *temp, int32_t *idx) {
    if (M <= 0) return; /* :100 */
    const int bs = ref_opt_n_threads(N);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float *ds = xyz + (size_t)b * N * 3;
        float *tp = temp + (size_t)b * N;
        int32_t *out = idx + (size_t)b * M;
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        int old = 0;
        out[0] = 0; /* :115-116 */
        for (int j = 1; j < M; ++j) {
            float x1 = ds[old * 3 + 0], y1 = ds[old * 3 + 1], z1 = ds[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) { /* one simulated thread at a time, :121-141 */
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < N; k += bs) {
                    float d = dist_direct(ds[k * 3 + 0] - x1, ds[k * 3 + 1] - y1, ds[k * 3 + 2] - z1);
                    float d2 = fminf(d, tp[k]);
                    tp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            /* the shared-memory tree, :143-203 with __update :86-91 */
            for (int s = bs / 2; s >= 1; s >>= 1)
                for (int tid = 0; tid < s; ++tid) {
                    float v1 = dists[tid], v2 = dists[tid + s];
                    int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2; /* max(v1,v2) */
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* ---------------------------------------------------------------------------------------- */
/* ball_query -- ball_query_gpu.cu:9-45                                                      */
/* ---------------------------------------------------------------------------------------- */
void orc_ball_query(int B, int N, int M, float radius, int nsample, const float *new_xyz,
                    const float *xyz, int32_t *idx) {
    const float radius2 = radius * radius; /* :24, FP32 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < M; ++p) {
            const float *c = new_xyz + ((size_t)b * M + p) * 3;
            const float *pts = xyz + (size_t)b * N * 3;
            int32_t *o = idx + ((size_t)b * M + p) * nsample;
            int cnt = 0;
            for (int k = 0; k < N; ++k) {
                float d2 = dist_direct(c[0] - pts[k * 3 + 0], c[1] - pts[k * 3 + 1],
                                       c[2] - pts[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
}

/* ---------------------------------------------------------------------------------------- */
/* three_nn / three_interpolate -- interpolate_gpu.cu                                        */
/* ---------------------------------------------------------------------------------------- */
void orc_three_nn(int B, int n, int m, const float *unknown, const float *known, float *dist2,
                  int32_t *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)b * n + p) * 3;
            const float *kn = known + (size_t)b * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40; /* :28 */
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = dist_direct(u[0] - kn[k * 3 + 0], u[1] - kn[k * 3 + 1], u[2] - kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; besti3 = besti2;
                    best2 = best1; besti2 = besti1;
                    best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2;
                    best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            float *od = dist2 + ((size_t)b * n + p) * 3;
            int32_t *oi = idx + ((size_t)b * n + p) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
        }
}

void orc_three_interpolate(int B, int C, int m, int n, const float *points, const int32_t *idx,
                           const float *weight, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *p = points + ((size_t)b * C + c) * m;
            for (int i = 0; i < n; ++i) {
                const float *w = weight + ((size_t)b * n + i) * 3;
                const int32_t *id = idx + ((size_t)b * n + i) * 3;
                /* :96 as compiled: FMUL w1*p1; FFMA w0*p0+.; FFMA w2*p2+. */
                out[((size_t)b * C + c) * n + i] =
                    fmaf(w[2], p[id[2]], fmaf(w[0], p[id[0]], w[1] * p[id[1]]));
            }
        }
}

void orc_three_interpolate_grad(int B, int C, int n, int m, const float *grad_out,
                                const int32_t *idx, const float *weight, float *grad_points) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            float *gp = grad_points + ((size_t)b * C + c) * m;
            for (int i = 0; i < n; ++i) {
                float g = grad_out[((size_t)b * C + c) * n + i];
                const float *w = weight + ((size_t)b * n + i) * 3;
                const int32_t *id = idx + ((size_t)b * n + i) * 3;
                gp[id[0]] += g * w[0];
                gp[id[1]] += g * w[1];
                gp[id[2]] += g * w[2];
            }
        }
}

/* ---------------------------------------------------------------------------------------- */
/* gather / group                                                                            */
/* ---------------------------------------------------------------------------------------- */
void orc_gather(int B, int C, int N, int M, const float *points, const int32_t *idx, float *out) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                out[((size_t)b * C + c) * M + j] =
                    points[((size_t)b * C + c) * N + idx[(size_t)b * M + j]];
}

void orc_gather_grad(int B, int C, int N, int M, const float *grad_out, const int32_t *idx,
                     float *grad_points) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                grad_points[((size_t)b * C + c) * N + idx[(size_t)b * M + j]] +=
                    grad_out[((size_t)b * C + c) * M + j];
}

void orc_group(int B, int C, int N, int np, int ns, const float *points, const int32_t *idx,
               float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *p = points + ((size_t)b * C + c) * N;
            float *o = out + ((size_t)b * C + c) * np * ns;
            const int32_t *id = idx + (size_t)b * np * ns;
            for (size_t t = 0; t < (size_t)np * ns; ++t) o[t] = p[id[t]];
        }
}

void orc_group_grad(int B, int C, int N, int np, int ns, const float *grad_out,
                    const int32_t *idx, float *grad_points) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            float *gp = grad_points + ((size_t)b * C + c) * N;
            const float *g = grad_out + ((size_t)b * C + c) * np * ns;
            const int32_t *id = idx + (size_t)b * np * ns;
            for (size_t t = 0; t < (size_t)np * ns; ++t) gp[id[t]] += g[t];
        }
}

/* ---------------------------------------------------------------------------------------- */
/* Chamfer -- pytorch3d.loss.chamfer_distance defaults (PARITY UNPINNED, see header)          */
/* ---------------------------------------------------------------------------------------- */
static void nn_dir(int N, int M, const float *x, const float *y, float *dx, int32_t *ix,
                   double *sum) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int i = 0; i < N; ++i) {
        float best = INFINITY;
        int bi = 0;
        for (int j = 0; j < M; ++j) {
            float d = dist_direct_xyz(x[i * 3 + 0] - y[j * 3 + 0], x[i * 3 + 1] - y[j * 3 + 1],
                                      x[i * 3 + 2] - y[j * 3 + 2]);
            if (d < best) {
                best = d;
                bi = j;
            }
        }
        if (dx) dx[i] = best;
        if (ix) ix[i] = bi;
        s += (double)best;
    }
    *sum = s;
}

double orc_chamfer(int B, int N, int M, const float *x, const float *y, float *dx, float *dy,
                   int32_t *ix, int32_t *iy) {
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        double sx, sy;
        nn_dir(N, M, x + (size_t)b * N * 3, y + (size_t)b * M * 3, dx ? dx + (size_t)b * N : NULL,
               ix ? ix + (size_t)b * N : NULL, &sx);
        nn_dir(M, N, y + (size_t)b * M * 3, x + (size_t)b * N * 3, dy ? dy + (size_t)b * M : NULL,
               iy ? iy + (size_t)b * M : NULL, &sy);
        total += sx / N + sy / M;
    }
    return total / B;
}

/* ---------------------------------------------------------------------------------------- */
/* EMD -- models/EMD/cuda/emd_kernel.cu                                                      */
/* ---------------------------------------------------------------------------------------- */
static inline float emd_w(float level, const float *a, const float *b) {
    /* :82-83: d = level * ((x2-x1)^2+...), w = __expf(d). a = xyz1 point, b = xyz2 point. */
    float d = level * dist_direct(b[0] - a[0], b[1] - a[1], b[2] - a[2]);
    return expf(d);
}

void orc_emd_approxmatch(int B, int n, int m, const float *xyz1, const float *xyz2, float *match) {
    float multiL, multiR;
    if (n >= m) { multiL = 1; multiR = (float)(n / m); } /* :33-38, integer division */
    else        { multiL = (float)(m / n); multiR = 1; }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < B; ++i) {
        const float *p1 = xyz1 + (size_t)i * n * 3, *p2 = xyz2 + (size_t)i * m * 3;
        float *mt = match + (size_t)i * n * m;
        float *remainL = (float *)malloc(sizeof(float) * n), *ratioL = (float *)malloc(sizeof(float) * n);
        float *remainR = (float *)malloc(sizeof(float) * m), *ratioR = (float *)malloc(sizeof(float) * m);
        memset(mt, 0, sizeof(float) * (size_t)n * m);
        for (int k = 0; k < n; ++k) remainL[k] = multiL;
        for (int l = 0; l < m; ++l) remainR[l] = multiR;
        for (int j = 7; j >= -2; --j) {
            float level = -powf(4.0f, (float)j);
            if (j == -2) level = 0;
            for (int k = 0; k < n; ++k) { /* :54-87 */
                float suml = 1e-9f;
                for (int l = 0; l < m; ++l)
                    suml = fmaf(emd_w(level, p1 + k * 3, p2 + l * 3), remainR[l], suml);
                ratioL[k] = remainL[k] / suml;
            }
            for (int l = 0; l < m; ++l) { /* :89-123 */
                float sumr = 0;
                for (int k = 0; k < n; ++k)
                    sumr = fmaf(emd_w(level, p1 + k * 3, p2 + l * 3), ratioL[k], sumr);
                sumr *= remainR[l];
                float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
            }
            for (int k = 0; k < n; ++k) { /* :125-158 */
                float suml = 0;
                float rl = ratioL[k];
                for (int l = 0; l < m; ++l) {
                    float t = rl * emd_w(level, p1 + k * 3, p2 + l * 3); /* FMUL rl*e */
                    mt[(size_t)l * n + k] = fmaf(t, ratioR[l], mt[(size_t)l * n + k]);
                    suml = fmaf(t, ratioR[l], suml);
                }
                remainL[k] = fmaxf(0.0f, remainL[k] - suml);
            }
        }
        free(remainL); free(ratioL); free(remainR); free(ratioR);
    }
}

void orc_emd_matchcost(int B, int n, int m, const float *xyz1, const float *xyz2,
                       const float *match, float *cost) {
    enum { T = 512 }; /* launch <<<32,512>>>, :278 */
    for (int i = 0; i < B; ++i) {
        const float *p1 = xyz1 + (size_t)i * n * 3, *p2 = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        float allsum[T];
#pragma omp parallel for schedule(static)
        for (int t = 0; t < T; ++t) {
            float subsum = 0;
            for (int k = t; k < n; k += T)
                for (int l = 0; l < m; ++l) {
                    float d = dist_direct(p2[l * 3 + 0] - p1[k * 3 + 0], p2[l * 3 + 1] - p1[k * 3 + 1],
                                          p2[l * 3 + 2] - p1[k * 3 + 2]);
                    subsum = fmaf(d, mt[(size_t)l * n + k], subsum);
                }
            allsum[t] = subsum;
        }
        for (int j = 1; j < T; j <<= 1) /* :236-241 */
            for (int t = 0; t < T; ++t)
                if ((t & j) == 0 && t + j < T && (t & (j - 1)) == 0) allsum[t] += allsum[t + j];
        cost[i] = allsum[0];
    }
}

void orc_emd_matchcost_grad(int B, int n, int m, const float *grad_cost, const float *xyz1,
                            const float *xyz2, const float *match, float *grad1, float *grad2) {
    for (int i = 0; i < B; ++i) {
        const float *p1 = xyz1 + (size_t)i * n * 3, *p2 = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        for (int l = 0; l < n; ++l) { /* matchcostgrad1 :337-359 */
            float dx = 0, dy = 0, dz = 0;
            for (int k = 0; k < m; ++k) {
                float d = mt[(size_t)k * n + l] * 2;
                dx = fmaf(p1[l * 3 + 0] - p2[k * 3 + 0], d, dx);
                dy = fmaf(p1[l * 3 + 1] - p2[k * 3 + 1], d, dy);
                dz = fmaf(p1[l * 3 + 2] - p2[k * 3 + 2], d, dz);
            }
            grad1[((size_t)i * n + l) * 3 + 0] = dx * grad_cost[i];
            grad1[((size_t)i * n + l) * 3 + 1] = dy * grad_cost[i];
            grad1[((size_t)i * n + l) * 3 + 2] = dz * grad_cost[i];
        }
        for (int k = 0; k < m; ++k) { /* matchcostgrad2 :290-331 (summation order simplified) */
            double sx = 0, sy = 0, sz = 0;
            for (int j = 0; j < n; ++j) {
                float d = mt[(size_t)k * n + j] * 2;
                sx += (double)((p2[k * 3 + 0] - p1[j * 3 + 0]) * d);
                sy += (double)((p2[k * 3 + 1] - p1[j * 3 + 1]) * d);
                sz += (double)((p2[k * 3 + 2] - p1[j * 3 + 2]) * d);
            }
            grad2[((size_t)i * m + k) * 3 + 0] = (float)sx * grad_cost[i];
            grad2[((size_t)i * m + k) * 3 + 1] = (float)sy * grad_cost[i];
            grad2[((size_t)i * m + k) * 3 + 2] = (float)sz * grad_cost[i];
        }
    }
}
