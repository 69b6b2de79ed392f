/* [isC,isF,As] = mis_set(A,theta) -- MEX replacement of the reference's AMG/mis_set.m:1-68
 * (theta defaults to 0.025, :9-11).  Consumes the library's MATLAB-compatible rand stream. */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 1 || !mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:nargin", "[isC,isF,As] = mis_set(A,theta)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t N = mxGetN(prhs[0]);
    ssn_csr A, As; memset(&A, 0, sizeof(A)); memset(&As, 0, sizeof(As));
    ssn_mex_upload_sparse(c, prhs[0], &A);
    uint8_t *flags = (uint8_t *)ssn_mex_dev_alloc(c, 2 * N);
    int st = ssn_mis_set(c, &A, nrhs > 1 ? mxGetScalar(prhs[1]) : 0.025, flags, flags + N, nlhs > 2 ? &As : NULL);
    plhs[0] = ssn_mex_logical_from_device(c, flags, N, &st);
    if (nlhs > 1) plhs[1] = ssn_mex_logical_from_device(c, flags + N, N, &st);
    if (nlhs > 2 && st == SSN_OK) plhs[2] = ssn_mex_csr_to_sparse(c, &As);      /* symmetric pattern */
    ssn_free(c, flags); ssn_csr_free(c, &A); ssn_csr_free(c, &As);
    ssn_mex_check(st);
}
