/* [blocks,sizes,p,r] = components(A) -- MEX replacement of the reference's components.m:1-64
 * (component order: ascending smallest member, members ascending). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 1 || !mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:nargin", "[blocks,sizes,p,r] = components(A)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t N = mxGetN(prhs[0]);
    ssn_csr A; memset(&A, 0, sizeof(A)); ssn_mex_upload_sparse(c, prhs[0], &A);
    int32_t *dev = (int32_t *)ssn_mex_dev_alloc(c, sizeof(int32_t) * (4 * N + 1));
    int32_t *blocks = dev, *sizes = dev + N, *p = dev + 2 * N, *r = dev + 3 * N;
    int nc = 0;
    int st = ssn_components(c, &A, blocks, sizes, p, r, &nc);
    int32_t *h = (int32_t *)mxMalloc(sizeof(int32_t) * (4 * N + 1));
    if (st == SSN_OK) st = ssn_memcpy_d2h(c, h, dev, sizeof(int32_t) * (4 * N + 1));
    if (st == SSN_OK) {
        plhs[0] = mxCreateDoubleMatrix(1, (mwSize)N, mxREAL);
        for (size_t i = 0; i < N; ++i) mxGetPr(plhs[0])[i] = h[i];
        if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(1, (mwSize)nc, mxREAL); for (int i = 0; i < nc; ++i) mxGetPr(plhs[1])[i] = h[N + i]; }
        if (nlhs > 2) { plhs[2] = mxCreateDoubleMatrix(1, (mwSize)N, mxREAL); for (size_t i = 0; i < N; ++i) mxGetPr(plhs[2])[i] = h[2 * N + i] + 1; }
        if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix(1, (mwSize)nc + 1, mxREAL); for (int i = 0; i <= nc; ++i) mxGetPr(plhs[3])[i] = h[3 * N + i] + 1; }
    }
    mxFree(h); ssn_free(c, dev); ssn_csr_free(c, &A);
    ssn_mex_check(st);
}
