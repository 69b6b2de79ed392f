/* [Ac,Pro,As,indC] = transfer(A,amg_options) -- MEX replacement of the reference's AMG/transfer.m:1-67.
 * The reference reads the level counter from `global J` (transfer.m:17); the shim reads the same
 * global through mexGetVariablePtr so that Class_AMG.m's own loop (if kept in MATLAB) still works. */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 1 || !mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:nargin", "[Ac,Pro,As,indC] = transfer(A,amg_options)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t N = mxGetN(prhs[0]);
    ssn_csr A, Ac, Pro, As; memset(&A, 0, sizeof(A)); memset(&Ac, 0, sizeof(Ac)); memset(&Pro, 0, sizeof(Pro)); memset(&As, 0, sizeof(As));
    ssn_mex_upload_sparse(c, prhs[0], &A);
    ssn_amg_options o; double *guess = NULL; ssn_mex_amg_options(c, nrhs > 1 ? prhs[1] : NULL, 0, &o, &guess);
    const mxArray *J = mexGetVariablePtr("global", "J");
    const int level = (J && !mxIsEmpty(J)) ? (int)mxGetScalar(J) : 2;                      /* no global J: general branch */
    uint8_t *indC = (uint8_t *)ssn_mex_dev_alloc(c, N);
    int st = ssn_transfer(c, &A, nrhs > 1 ? &o : NULL, level, &Ac, &Pro, nlhs > 2 ? &As : NULL, indC);
    if (st == SSN_OK) plhs[0] = ssn_mex_csr_to_sparse(c, &Ac);                               /* symmetric pattern */
    if (nlhs > 1) plhs[1] = ssn_mex_csr_to_sparse_general(c, &Pro, &st);
    if (nlhs > 2 && st == SSN_OK) plhs[2] = ssn_mex_csr_to_sparse(c, &As);
    if (nlhs > 3) plhs[3] = ssn_mex_logical_from_device(c, indC, N, &st);
    ssn_free(c, indC); ssn_csr_free(c, &A); ssn_csr_free(c, &Ac); ssn_csr_free(c, &Pro); ssn_csr_free(c, &As);
    ssn_mex_check(st);
}
