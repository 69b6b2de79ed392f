/* y = ASAtz(z,s,p,q) -- MEX replacement of the reference's ASAtz.m:2-23 (as written: Q*p at :21). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    (void)nlhs;
    if (nrhs != 4 || !mxIsLogical(prhs[1])) mexErrMsgIdAndTxt("ssnamg:nargin", "y = ASAtz(z,s,p,q)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[2]), n = mxGetNumberOfElements(prhs[3]), N = m + n;
    double *z = ssn_mex_to_device(c, prhs[0], N), *p = ssn_mex_to_device(c, prhs[2], m), *q = ssn_mex_to_device(c, prhs[3], n);
    uint8_t *s = (uint8_t *)ssn_mex_dev_alloc(c, m * n);
    double *y = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    int st = ssn_memcpy_h2d(c, s, mxGetLogicals(prhs[1]), m * n);
    if (st == SSN_OK) st = ssn_asatz(c, z, s, p, q, (int64_t)m, (int64_t)n, y);
    plhs[0] = ssn_mex_from_device(c, y, N, &st);
    ssn_free(c, z); ssn_free(c, p); ssn_free(c, q); ssn_free(c, s); ssn_free(c, y);
    ssn_mex_check(st);
}
