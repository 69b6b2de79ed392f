/* [x,it,rel_res,rel_resk,rhok] = twogrid(A,b,amg_options) -- MEX replacement of the reference's
 * AMG/twogrid.m:1-150.  The function's own defaults are applied here (:16-34): with two arguments
 * retol 1e-12, bigph 0, maxit 20, smoth 10, isnsp 1; empty fields of a given struct -> retol 0, bigph 0,
 * maxit 50, smoth 3, isnsp 0, fnode 0 (the library would otherwise fall back to Class_AMG's). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 2 || !mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:nargin", "[x,it,rel_res,rel_resk,rhok] = twogrid(A,b,amg_options)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t N = mxGetN(prhs[0]);
    ssn_csr A; memset(&A, 0, sizeof(A)); ssn_mex_upload_sparse(c, prhs[0], &A);
    double *b = ssn_mex_to_device(c, prhs[1], N), *guess = NULL;
    ssn_amg_options o; ssn_mex_amg_options(c, nrhs > 2 ? prhs[2] : NULL, N, &o, &guess);
    if (nrhs == 2) { o.retol = 1e-12; o.bigph = 0; o.maxit = 20; o.smoth = 10; o.isnsp = 1; }   /* :16-21 */
    if (o.retol < 0) o.retol = 0.0;                                                        /* :22-34 */
    if (o.bigph < 0) o.bigph = 0;
    if (o.maxit < 0) o.maxit = 50;
    if (o.smoth < 0) o.smoth = 3;
    if (o.isnsp < 0) o.isnsp = 0;
    const int maxit = o.maxit;
    double *x = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    double *relk = (double *)mxCalloc((size_t)maxit + 2, sizeof(double)), *rho = (double *)mxCalloc((size_t)maxit + 2, sizeof(double));
    int it = 0, len = 0; double rel = 0;
    int st = ssn_twogrid(c, &A, b, &o, x, &it, &rel, relk, rho, &len);
    plhs[0] = ssn_mex_from_device(c, x, N, &st);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(it);
    if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(rel);
    if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix((mwSize)len, 1, mxREAL); memcpy(mxGetPr(plhs[3]), relk, sizeof(double) * (size_t)len); }
    if (nlhs > 4) { plhs[4] = mxCreateDoubleMatrix((mwSize)len, 1, mxREAL); memcpy(mxGetPr(plhs[4]), rho, sizeof(double) * (size_t)len); }
    mxFree(relk); mxFree(rho);
    ssn_free(c, b); ssn_free(c, x); if (guess) ssn_free(c, guess); ssn_csr_free(c, &A);
    ssn_mex_check(st);
}
