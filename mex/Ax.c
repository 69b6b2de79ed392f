/* y = Ax(x,p,q)  -- MEX replacement of the reference's Ax.m:2-14.
 * Build (where MATLAB exists):  mex -R2018a Ax.c -L.. -lssnamg  */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 3) mexErrMsgIdAndTxt("ssnamg:nargin", "y = Ax(x,p,q)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[1]), n = mxGetNumberOfElements(prhs[2]);
    plhs[0] = mxCreateDoubleMatrix((mwSize)(n + m), 1, mxREAL);
    const mxArray *x = prhs[0];
    mxArray *xfull = NULL;
    if (mxIsSparse(x)) {                       /* Class1/warmup_class1.m:29 passes a sparse zero x */
        mxArray *in = (mxArray *)x;
        mexCallMATLAB(1, &xfull, 1, &in, "full");
        x = xfull;
    }
    if (mxGetNumberOfElements(x) != m * n) mexErrMsgIdAndTxt("ssnamg:arg", "numel(x) must be m*n");
    /* host mxArrays: the _host entry point does the H2D/D2H copies itself */
    int st = ssn_ax_host(c, mxGetPr(x), mxGetPr(prhs[1]), mxGetPr(prhs[2]), (int64_t)m, (int64_t)n, mxGetPr(plhs[0]));
    if (xfull) mxDestroyArray(xfull);
    ssn_mex_check(st);
}
