/* [uk,lk] = warmup_class2(c,r,l,p,q,mu,phi,res,maxit) -- MEX replacement of the reference's
 * Class2/warmup_class2.m:2-108 (fused A-ADMM kernels of partial OT, ssn_warmup_class2).  nargin rules of :3-18. */
#include <math.h>
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 7) mexErrMsgIdAndTxt("ssnamg:nargin", "[uk,lk] = warmup_class2(c,r,l,p,q,mu,phi,res,maxit)");
    ssn_ctx *ctx = ssn_mex_ctx();
    const size_t n = mxGetNumberOfElements(prhs[1]), m = mxGetNumberOfElements(prhs[2]), N = m + n;
    double res = nrhs > 7 ? mxGetScalar(prhs[7]) : 1e-1, maxit = nrhs > 8 ? mxGetScalar(prhs[8]) : INFINITY;
    if (nrhs == 9 && res == 0 && isinf(maxit)) mexErrMsgIdAndTxt("ssnamg:arg", "res = 0 and maxit = inf");       /* :10-12 */
    if (isinf(maxit)) maxit = 500;                                                                               /* :16-18 */
    double *c = ssn_mex_to_device(ctx, prhs[0], m * n), *p = ssn_mex_to_device(ctx, prhs[3], m), *q = ssn_mex_to_device(ctx, prhs[4], n);
    double *phi = ssn_mex_to_device(ctx, prhs[6], m * n);
    double *b = (double *)ssn_mex_dev_alloc(ctx, (N + 1) * sizeof(double));                                      /* b = [r;l;mu], :21 */
    const double mu = mxGetScalar(prhs[5]);
    ssn_mex_check(ssn_memcpy_h2d(ctx, b, mxGetPr(prhs[1]), n * sizeof(double)));
    ssn_mex_check(ssn_memcpy_h2d(ctx, b + n, mxGetPr(prhs[2]), m * sizeof(double)));
    ssn_mex_check(ssn_memcpy_h2d(ctx, b + N, &mu, sizeof(double)));
    double *uk = (double *)ssn_mex_dev_alloc(ctx, (m * n + N) * sizeof(double)), *lk = (double *)ssn_mex_dev_alloc(ctx, (N + 1) * sizeof(double));
    int st = ssn_warmup_class2(ctx, c, b, p, q, (int64_t)m, (int64_t)n, phi, (int)maxit, uk, lk);
    plhs[0] = ssn_mex_from_device(ctx, uk, m * n + N, &st);
    if (nlhs > 1) plhs[1] = ssn_mex_from_device(ctx, lk, N + 1, &st);
    ssn_free(ctx, c); ssn_free(ctx, p); ssn_free(ctx, q); ssn_free(ctx, phi); ssn_free(ctx, b); ssn_free(ctx, uk); ssn_free(ctx, lk);
    ssn_mex_check(st);
}
