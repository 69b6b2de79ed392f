/* y = invAAt(x,p,q,sg1,sg2) -- MEX replacement of the reference's invAAt.m:1-21 (nargin defaults :7-12). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    (void)nlhs;
    if (nrhs < 3) mexErrMsgIdAndTxt("ssnamg:nargin", "y = invAAt(x,p,q,sg1,sg2)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[1]), n = mxGetNumberOfElements(prhs[2]), N = m + n;
    const double sg1 = nrhs > 3 ? mxGetScalar(prhs[3]) : 1.0, sg2 = nrhs > 4 ? mxGetScalar(prhs[4]) : sg1;
    double *x = ssn_mex_to_device(c, prhs[0], N), *p = ssn_mex_to_device(c, prhs[1], m), *q = ssn_mex_to_device(c, prhs[2], n);
    double *y = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    int st = ssn_invaat(c, x, p, q, (int64_t)m, (int64_t)n, sg1, sg2, y);
    plhs[0] = ssn_mex_from_device(c, y, N, &st);
    ssn_free(c, x); ssn_free(c, p); ssn_free(c, q); ssn_free(c, y);
    ssn_mex_check(st);
}
