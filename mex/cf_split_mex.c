/* [indC,indF] = cf_split_mex(S) -- the computational part of the reference's AMG/cf_split.m:1-16.
 * The third output of cf_split is a MATLAB graph object, which C cannot build; the drop-in is this
 * MEX plus the two-line wrapper
 *     function [indC,indF,SubG] = cf_split(S)
 *     [indC,indF] = cf_split_mex(S); SubG = graph(S);           % cf_split.m:6
 * saved as cf_split.m in the mex directory. */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 1 || !mxIsSparse(prhs[0])) mexErrMsgIdAndTxt("ssnamg:nargin", "[indC,indF] = cf_split_mex(S)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t N = mxGetN(prhs[0]);
    ssn_csr S; memset(&S, 0, sizeof(S)); ssn_mex_upload_sparse(c, prhs[0], &S);
    uint8_t *flags = (uint8_t *)ssn_mex_dev_alloc(c, 2 * N);
    int st = ssn_cf_split(c, &S, flags, flags + N);
    plhs[0] = ssn_mex_logical_from_device(c, flags, N, &st);
    if (nlhs > 1) plhs[1] = ssn_mex_logical_from_device(c, flags + N, N, &st);
    ssn_free(c, flags); ssn_csr_free(c, &S);
    ssn_mex_check(st);
}
