// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product library.
//
// extern "C" harness around the *unmodified* reference launchers so the GPU parity tests can
// call the reference's own kernels (compiled from /root/reference/pointnet2/src/*.cu where
// they lie, see oracle/Makefile) through ctypes with raw device pointers. The launcher
// prototypes are the ones declared in the reference's *_gpu.h headers
// (sampling_gpu.h:12-27, ball_query_gpu.h:12-13, group_points_gpu.h:13-20,
// interpolate_gpu.h:13-28); they are re-declared here so this file needs no torch headers.
#include <cuda_runtime.h>

void furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset,
                                             float *temp, int *idxs, cudaStream_t stream);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float *points,
                                        const int *idx, float *out, cudaStream_t stream);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints,
                                             const float *grad_out, const int *idx,
                                             float *grad_points, cudaStream_t stream);
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                     const float *new_xyz, const float *xyz, int *idx,
                                     cudaStream_t stream);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                       const float *points, const int *idx, float *out,
                                       cudaStream_t stream);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                            const float *grad_out, const int *idx,
                                            float *grad_points, cudaStream_t stream);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown, const float *known,
                                   float *dist2, int *idx, cudaStream_t stream);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float *points,
                                            const int *idx, const float *weight, float *out,
                                            cudaStream_t stream);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m,
                                                 const float *grad_out, const int *idx,
                                                 const float *weight, float *grad_points,
                                                 cudaStream_t stream);

extern "C" {
void ref_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, void *s) {
    furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx, (cudaStream_t)s);
}
void ref_gather(int b, int c, int n, int np, const float *p, const int *idx, float *out, void *s) {
    gather_points_kernel_launcher_fast(b, c, n, np, p, idx, out, (cudaStream_t)s);
}
void ref_gather_grad(int b, int c, int n, int np, const float *g, const int *idx, float *gp, void *s) {
    gather_points_grad_kernel_launcher_fast(b, c, n, np, g, idx, gp, (cudaStream_t)s);
}
void ref_ball_query(int b, int n, int m, float r, int ns, const float *new_xyz, const float *xyz,
                    int *idx, void *s) {
    ball_query_kernel_launcher_fast(b, n, m, r, ns, new_xyz, xyz, idx, (cudaStream_t)s);
}
void ref_group(int b, int c, int n, int np, int ns, const float *p, const int *idx, float *out, void *s) {
    group_points_kernel_launcher_fast(b, c, n, np, ns, p, idx, out, (cudaStream_t)s);
}
void ref_group_grad(int b, int c, int n, int np, int ns, const float *g, const int *idx, float *gp, void *s) {
    group_points_grad_kernel_launcher_fast(b, c, n, np, ns, g, idx, gp, (cudaStream_t)s);
}
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                  int *idx, void *s) {
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, (cudaStream_t)s);
}
void ref_three_interpolate(int b, int c, int m, int n, const float *p, const int *idx,
                           const float *w, float *out, void *s) {
    three_interpolate_kernel_launcher_fast(b, c, m, n, p, idx, w, out, (cudaStream_t)s);
}
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *g, const int *idx,
                                const float *w, float *gp, void *s) {
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, g, idx, w, gp, (cudaStream_t)s);
}
}
