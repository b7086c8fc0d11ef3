/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of MoCoPCI's point-set neighbourhood
 * hot path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (mocopci_b200/libb200pci.so) never does.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Floating point: compiled with -ffp-contract=off; every fused multiply-add the reference's
 * compiled kernels perform (checked in their SASS, see DESIGN.md "FP contraction") is written
 * as an explicit fmaf().
 *
 * Pinning status (see DESIGN.md section "Oracle"):
 *   - orc_knn_expanded:  pinned against the reference's own torch code imported from
 *     /root/reference/models/pointconv_util.py (tests/golden/knn_*.npz, made by
 *     tests/golden/make_golden.py) -- distances bitwise, index sets per SURVEY section 8c.
 *   - orc_emd_*:         pinned against the reference's only known-answer vector
 *     (models/EMD/test_emd_loss.py:7-18).
 *   - orc_fps / ball_query / three_nn / three_interpolate / gather / group: pinned on the GPU box
 *     against the reference's own CUDA kernels compiled unmodified into oracle/_ref/ (bitwise).
 *   - orc_chamfer:       PARITY UNPINNED -- pytorch3d 0.7.5 (environment.yaml:90) is not vendored
 *     and not installed; restates its documented semantics.
 */
#ifndef MOCOPCI_ORACLE_H
#define MOCOPCI_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* models/pointconv_util.py:67-88 square_distance(src=query, dst=ref) -> D[b,s,n].
 * q: [B,S,3], r: [B,N,3] contiguous. out: [B,S,N]. */
void orc_square_distance(int B, int S, int N, const float *q, const float *r, float *out);

/* models/pointconv_util.py:129-140 knn_point(k, xyz=r, new_xyz=q): k smallest of D per query.
 * Output sorted ascending by (distance, index) -- lowest index wins ties (north_star).
 * idx: int64 [B,S,k]; dist (optional, may be NULL): [B,S,k]. Returns 0, or -1 if k > N. */
int orc_knn_expanded(int B, int S, int N, int k, const float *q, const float *r, int64_t *idx,
                     float *dist);

/* Direct-difference k-NN with the three_nn arithmetic (interpolate_gpu.cu:37), any k.
 * Sorted ascending by (distance, index). Slots beyond N are (inf, 0) like three_nn's m<3 case. */
int orc_knn_direct(int B, int S, int N, int k, const float *q, const float *r, int64_t *idx,
                   float *dist);

/* The same selection in any of the four distance forms the hot path uses:
 * 0 expanded (torch square_distance), 1 pointnet2 kernels fma(dz,dz,fma(dx,dx,dy*dy)),
 * 2 pytorch3d CUDA knn fma(dz,dz,fma(dy,dy,dx*dx)) (restated from its published source, unpinned),
 * 3 models/pointT_layer2.py:20 (dx*dx + dy*dy) + dz*dz without FMA (CPU torch's sum order),
 * 4 the same as CUDA torch sums it, (dx*dx + dz*dz) + dy*dy, 5 form 0 with the norms summed in
 * CUDA torch's order (x*x + z*z) + y*y. */
int orc_knn_form(int form, int B, int S, int N, int k, const float *q, const float *r, int64_t *idx,
                 float *dist);

/* models/pointconv_util.py:111-127,142-153 (f2): k smallest cosine distances 1 - <q/|q|, r/|r|>,
 * q [B,S,C], r [B,N,C]; dot products in double (the reference's sgemm order is unspecified:
 * compare with a tolerance). Sorted by (distance, index). */
int orc_knn_cosine(int B, int S, int N, int C, int k, const float *q, const float *r, int64_t *idx,
                   float *dist);

/* pointnet2/pointnet2_modules.py:139-144 + pointnet2_utils.py:97: dist2 [rows,3] -> dist = sqrt,
 * weight = (1/(dist+eps)) / sum. */
void orc_three_nn_weights(long long rows, float eps, const float *dist2, float *dist, float *weight);

/* pointnet2/src/sampling_gpu.cu:93-209 + cuda_utils.h:10-14. xyz [B,N,3], temp [B,N] in/out
 * (caller pre-fills 1e10, pointnet2_utils.py:26), idx int32 [B,M]. */
void orc_fps(int B, int N, int M, const float *xyz, float *temp, int32_t *idx);

/* pointnet2/src/ball_query_gpu.cu:9-45. idx int32 [B,M,nsample] must be pre-zeroed by the
 * caller (pointnet2_utils.py:218). */
void orc_ball_query(int B, int N, int M, float radius, int nsample, const float *new_xyz,
                    const float *xyz, int32_t *idx);

/* pointnet2/src/interpolate_gpu.cu:9-52. dist2 [B,n,3] (squared, the kernel's output before
 * pointnet2_utils.py:99 takes sqrt), idx int32 [B,n,3]. */
void orc_three_nn(int B, int n, int m, const float *unknown, const float *known, float *dist2,
                  int32_t *idx);

/* pointnet2/src/interpolate_gpu.cu:77-97. points [B,C,m], idx [B,n,3], weight [B,n,3] -> out [B,C,n]. */
void orc_three_interpolate(int B, int C, int m, int n, const float *points, const int32_t *idx,
                           const float *weight, float *out);
/* interpolate_gpu.cu:120-142. grad_points [B,C,m] must be pre-zeroed. */
void orc_three_interpolate_grad(int B, int C, int n, int m, const float *grad_out,
                                const int32_t *idx, const float *weight, float *grad_points);

/* sampling_gpu.cu:8-24 / :46-63. points [B,C,N], idx [B,M] -> out [B,C,M]. */
void orc_gather(int B, int C, int N, int M, const float *points, const int32_t *idx, float *out);
void orc_gather_grad(int B, int C, int N, int M, const float *grad_out, const int32_t *idx,
                     float *grad_points);

/* group_points_gpu.cu:47-66 / :8-25. points [B,C,N], idx [B,np,ns] -> out [B,C,np,ns]. */
void orc_group(int B, int C, int N, int np, int ns, const float *points, const int32_t *idx,
               float *out);
void orc_group_grad(int B, int C, int N, int np, int ns, const float *grad_out,
                    const int32_t *idx, float *grad_points);

/* models/utils.py:36-45 -> pytorch3d.loss.chamfer_distance defaults (PARITY UNPINNED; distances in
 * pytorch3d's CUDA order, form 2 above).
 * x [B,N,3], y [B,M,3]. Outputs (each may be NULL): per-point squared NN distances dx [B,N],
 * dy [B,M], NN indices ix [B,N], iy [B,M]. Returns the scalar loss
 * mean_b( mean_i dx + mean_j dy ), accumulated in double. */
double orc_chamfer(int B, int N, int M, const float *x, const float *y, float *dx, float *dy,
                   int32_t *ix, int32_t *iy);

/* models/EMD/cuda/emd_kernel.cu:29-162. xyz1 [B,n,3], xyz2 [B,m,3] -> match [B,m,n].
 * exp is libm expf (the GPU uses ex2.approx), so GPU comparisons use a tolerance. */
void orc_emd_approxmatch(int B, int n, int m, const float *xyz1, const float *xyz2, float *match);
/* emd_kernel.cu:204-247 (per-thread sequential partial sums + the :236-241 tree, blockDim 512). */
void orc_emd_matchcost(int B, int n, int m, const float *xyz1, const float *xyz2,
                       const float *match, float *cost);
/* emd_kernel.cu:290-359. */
void orc_emd_matchcost_grad(int B, int n, int m, const float *grad_cost, const float *xyz1,
                            const float *xyz2, const float *match, float *grad1, float *grad2);

#ifdef __cplusplus
}
#endif
#endif
