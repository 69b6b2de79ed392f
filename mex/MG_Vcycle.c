/* e = MG_Vcycle(r,isnsp,k) -- MEX replacement of the reference's AMG/MG_Vcycle.m:2-46, on the hierarchy a previous
 * Class_AMG setup (ssn_amg_setup) left in the context (the reference's globals Ack/Prok/Rk/J). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    (void)nlhs;
    if (nrhs < 1) mexErrMsgIdAndTxt("ssnamg:nargin", "e = MG_Vcycle(r,isnsp,k)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t n = mxGetNumberOfElements(prhs[0]);
    const int isnsp = nrhs > 1 ? (int)mxGetScalar(prhs[1]) : 0, k = nrhs > 2 ? (int)mxGetScalar(prhs[2]) : 1;   /* defaults :5-7 */
    double *r = ssn_mex_to_device(c, prhs[0], n);
    double *e = (double *)ssn_mex_dev_alloc(c, n * sizeof(double));
    double *zero = (double *)mxCalloc(n ? n : 1, sizeof(double));
    ssn_mex_check(ssn_memcpy_h2d(c, e, zero, n * sizeof(double))); mxFree(zero);
    int st = ssn_mg_vcycle(c, r, isnsp, k, e);
    plhs[0] = ssn_mex_from_device(c, e, n, &st);
    ssn_free(c, r); ssn_free(c, e);
    ssn_mex_check(st);
}
