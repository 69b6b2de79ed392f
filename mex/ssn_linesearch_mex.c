/* [lk_new,ll,cFk_new] = ssn_linesearch_mex(wk,lk_old,zeta,wlk,p,q,tk,bk1,gama,nu,delta,ll_max,cFk_old,ress)
 * -- the Armijo loop of Class1/APD_SsN_Class1.m:189-211 in one call (batch = 0: the adaptive schedule,
 * 64-128 backtracking steps per read of wk through the screened kernels while the trial plans are sparse,
 * 8 through the dense kernel otherwise).  Optional: the unmodified script keeps working with its own loop
 * over Aty/prox/norm. */
#include <math.h>
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 14) mexErrMsgIdAndTxt("ssnamg:nargin", "ssn_linesearch_mex: 14 inputs");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[4]), n = mxGetNumberOfElements(prhs[5]), N = m + n;
    double *w = ssn_mex_to_device(c, prhs[0], m * n), *lo = ssn_mex_to_device(c, prhs[1], N), *ze = ssn_mex_to_device(c, prhs[2], N);
    double *wl = ssn_mex_to_device(c, prhs[3], N), *p = ssn_mex_to_device(c, prhs[4], m), *q = ssn_mex_to_device(c, prhs[5], n);
    const int scalar_gama = mxGetNumberOfElements(prhs[8]) == 1;
    double *gama = scalar_gama ? NULL : ssn_mex_to_device(c, prhs[8], m * n);
    double *out = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    int ll = 0, passes = 0; double n2 = 0, cF = 0;
    int st = ssn_linesearch(c, w, lo, ze, wl, p, q, (int64_t)m, (int64_t)n, mxGetScalar(prhs[6]), mxGetScalar(prhs[7]), gama,
                            scalar_gama ? mxGetScalar(prhs[8]) : INFINITY, mxGetScalar(prhs[9]), mxGetScalar(prhs[10]),
                            (int)mxGetScalar(prhs[11]), mxGetScalar(prhs[12]), mxGetScalar(prhs[13]), 0, out, &ll, &n2, &cF, &passes);
    plhs[0] = ssn_mex_from_device(c, out, N, &st);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(ll);
    if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(cF);
    ssn_free(c, w); ssn_free(c, lo); ssn_free(c, ze); ssn_free(c, wl); ssn_free(c, p); ssn_free(c, q); if (gama) ssn_free(c, gama); ssn_free(c, out);
    ssn_mex_check(st);
}
