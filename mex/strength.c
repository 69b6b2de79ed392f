/* S = strength(A,which) -- MEX replacement of the reference's AMG/strength.m:1-19 (which defaults to 2). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    (void)nlhs;
    if (nrhs < 1 || !mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:nargin", "S = strength(A,which)");
    ssn_ctx *c = ssn_mex_ctx();
    ssn_csr A, S; memset(&A, 0, sizeof(A)); memset(&S, 0, sizeof(S));
    ssn_mex_upload_sparse(c, prhs[0], &A);
    int st = ssn_strength(c, &A, nrhs > 1 ? (int)mxGetScalar(prhs[1]) : 2, &S);
    plhs[0] = ssn_mex_csr_to_sparse_general(c, &S, &st);       /* which = 1 is not symmetric */
    ssn_csr_free(c, &A); ssn_csr_free(c, &S);
    ssn_mex_check(st);
}
