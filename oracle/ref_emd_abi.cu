// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product library.
//
// Raw-pointer extern "C" harness around the reference's EMD kernels. The reference file is
// #included where it lies (never copied); we launch its __global__ templates with exactly the
// launch configuration its host functions use (emd_kernel.cu:192, :278, :396-397):
// approxmatch/matchcost/matchcostgrad1 <<<32,512>>>, matchcostgrad2 <<<dim3(32,32),256>>>,
// all on the legacy default stream.
#include REF_EMD_KERNEL_CU

extern "C" {
// match: [b,m,n] zero-filled by the kernel itself; temp: [b,(n+m)*2] scratch.
int ref_emd_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match,
                        float *temp) {
    approxmatch<float><<<32, 512>>>(b, n, m, xyz1, xyz2, match, temp);
    return (int)cudaGetLastError();
}
int ref_emd_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2,
                      const float *match, float *cost) {
    matchcost<float><<<32, 512>>>(b, n, m, xyz1, xyz2, match, cost);
    return (int)cudaGetLastError();
}
int ref_emd_matchcost_grad(int b, int n, int m, const float *grad_cost, const float *xyz1,
                           const float *xyz2, const float *match, float *grad1, float *grad2) {
    matchcostgrad1<float><<<32, 512>>>(b, n, m, grad_cost, xyz1, xyz2, match, grad1);
    matchcostgrad2<float><<<dim3(32, 32), 256>>>(b, n, m, grad_cost, xyz1, xyz2, match, grad2);
    return (int)cudaGetLastError();
}
}
