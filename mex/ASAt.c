/* H = ASAt(s,p,q)  -- MEX replacement of the reference's ASAt.m:2-20.  s is the full logical
 * m*n vector built at Class1/APD_SsN_Class1.m:140 (one byte per entry). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 3) mexErrMsgIdAndTxt("ssnamg:nargin", "H = ASAt(s,p,q)");
    if (!mxIsLogical(prhs[0]) || mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:arg", "s must be a full logical vector");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[1]), n = mxGetNumberOfElements(prhs[2]);
    if (mxGetNumberOfElements(prhs[0]) != m * n) mexErrMsgIdAndTxt("ssnamg:arg", "numel(s) must be m*n");
    ssn_csr H; memset(&H, 0, sizeof(H));
    int st = ssn_asat_host(c, (const uint8_t *)mxGetLogicals(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]),
                           (int64_t)m, (int64_t)n, &H);
    if (st == SSN_OK) plhs[0] = ssn_mex_csr_to_sparse(c, &H);   /* symmetric pattern: CSR arrays == CSC arrays */
    ssn_csr_free(c, &H);
    ssn_mex_check(st);
}
