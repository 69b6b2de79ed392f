/* [xk,lk] = warmup_class1(c,r,l,p,q,gama,res,maxit) -- MEX replacement of the reference's
 * Class1/warmup_class1.m:2-96 (fused A-ADMM kernels, ssn_warmup_class1).  nargin rules of :3-20. */
#include <math.h>
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 6) mexErrMsgIdAndTxt("ssnamg:nargin", "[xk,lk] = warmup_class1(c,r,l,p,q,gama,res,maxit)");
    ssn_ctx *ctx = ssn_mex_ctx();
    const size_t n = mxGetNumberOfElements(prhs[1]), m = mxGetNumberOfElements(prhs[2]), N = m + n;
    double res = nrhs > 6 ? mxGetScalar(prhs[6]) : 1e-1, maxit = nrhs > 7 ? mxGetScalar(prhs[7]) : INFINITY;
    if (nrhs == 8 && res == 0 && isinf(maxit)) mexErrMsgIdAndTxt("ssnamg:arg", "res = 0 and maxit = inf");       /* :10-12 */
    if (isinf(maxit)) maxit = 500;                                                                               /* :18-20 */
    double *c = ssn_mex_to_device(ctx, prhs[0], m * n), *p = ssn_mex_to_device(ctx, prhs[3], m), *q = ssn_mex_to_device(ctx, prhs[4], n);
    double *b = (double *)ssn_mex_dev_alloc(ctx, N * sizeof(double));                                            /* b = [r;l], :24 */
    ssn_mex_check(ssn_memcpy_h2d(ctx, b, mxGetPr(prhs[1]), n * sizeof(double)));
    ssn_mex_check(ssn_memcpy_h2d(ctx, b + n, mxGetPr(prhs[2]), m * sizeof(double)));
    const int scalar_gama = mxGetNumberOfElements(prhs[5]) == 1;
    double *gama = scalar_gama ? NULL : ssn_mex_to_device(ctx, prhs[5], m * n);
    double *xk = (double *)ssn_mex_dev_alloc(ctx, m * n * sizeof(double)), *lk = (double *)ssn_mex_dev_alloc(ctx, N * sizeof(double));
    int st = ssn_warmup_class1(ctx, c, b, p, q, (int64_t)m, (int64_t)n, gama, scalar_gama ? mxGetScalar(prhs[5]) : INFINITY, (int)maxit, xk, lk);
    plhs[0] = ssn_mex_from_device(ctx, xk, m * n, &st);
    if (nlhs > 1) plhs[1] = ssn_mex_from_device(ctx, lk, N, &st);
    ssn_free(ctx, c); ssn_free(ctx, p); ssn_free(ctx, q); ssn_free(ctx, b); if (gama) ssn_free(ctx, gama); ssn_free(ctx, xk); ssn_free(ctx, lk);
    ssn_mex_check(st);
}
