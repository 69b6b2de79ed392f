/* [d,it,res,resk] = PCG(H,e,pcg_options) -- MEX replacement of the reference's PCG.m:1-105
 * (precd 1, 2, 5; nargin == 2 takes the defaults of PCG.m:18-23). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 2 || !mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:nargin", "[d,it,res,resk] = PCG(H,e,pcg_options)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t N = mxGetN(prhs[0]);
    ssn_csr H; memset(&H, 0, sizeof(H)); ssn_mex_upload_sparse(c, prhs[0], &H);
    double *e = ssn_mex_to_device(c, prhs[1], N), *guess = NULL;
    ssn_pcg_options o; ssn_mex_pcg_options(c, nrhs > 2 ? prhs[2] : NULL, N, &o, &guess);
    const int maxit = o.maxit >= 0 ? o.maxit : 10000;
    double *d = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    mxArray *resk = mxCreateDoubleMatrix((mwSize)(maxit > 0 ? maxit : 1), 1, mxREAL);      /* resk = zeros(maxit,1), PCG.m:74 */
    int it = 0; double res = 0;
    int st = ssn_pcg(c, &H, e, nrhs > 2 ? &o : NULL, d, &it, &res, mxGetPr(resk));
    plhs[0] = ssn_mex_from_device(c, d, N, &st);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(it);
    if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(res);
    if (nlhs > 3) plhs[3] = resk;
    ssn_free(c, e); ssn_free(c, d); if (guess) ssn_free(c, guess); ssn_csr_free(c, &H);
    ssn_mex_check(st);
}
