/* z = Aty(y,p,q)  -- MEX replacement of the reference's Aty.m:2-14. */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 3) mexErrMsgIdAndTxt("ssnamg:nargin", "z = Aty(y,p,q)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[1]), n = mxGetNumberOfElements(prhs[2]);
    if (mxGetNumberOfElements(prhs[0]) < n + m) mexErrMsgIdAndTxt("ssnamg:arg", "numel(y) must be n+m");
    plhs[0] = mxCreateDoubleMatrix((mwSize)(m * n), 1, mxREAL);
    ssn_mex_check(ssn_aty_host(c, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), (int64_t)m, (int64_t)n,
                               mxGetPr(plhs[0])));
}
