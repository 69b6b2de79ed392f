/* y = invHHt(v,p,q,sg,phi) -- MEX replacement of the reference's Class2/invHHt.m:1-18. */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    (void)nlhs;
    if (nrhs != 5) mexErrMsgIdAndTxt("ssnamg:nargin", "y = invHHt(v,p,q,sg,phi)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[1]), n = mxGetNumberOfElements(prhs[2]), N = m + n;
    double *v = ssn_mex_to_device(c, prhs[0], N + 1), *p = ssn_mex_to_device(c, prhs[1], m), *q = ssn_mex_to_device(c, prhs[2], n);
    double *phi = ssn_mex_to_device(c, prhs[4], m * n);
    double *y = (double *)ssn_mex_dev_alloc(c, (N + 1) * sizeof(double));
    int st = ssn_invhht(c, v, p, q, (int64_t)m, (int64_t)n, mxGetScalar(prhs[3]), phi, y);
    plhs[0] = ssn_mex_from_device(c, y, N + 1, &st);
    ssn_free(c, v); ssn_free(c, p); ssn_free(c, q); ssn_free(c, phi); ssn_free(c, y);
    ssn_mex_check(st);
}
