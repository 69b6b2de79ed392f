/* Test-only shim built like a .mex file (own copy of ssn_mex_common.h's static handle): [x,it] = keep(A,b,amg_options)
 * runs Class_AMG's setup + solve and KEEPS the hierarchy (the state the reference holds in `global Ack Prok J
 * smoth_it Rk` while Class_AMG runs, AMG/Class_AMG.m:43), so that another shim can use it.  A second entry point
 * reports this shim's view of the shared context. */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 3) mexErrMsgIdAndTxt("ssnamg:nargin", "[x,it] = keep(A,b,amg_options)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t N = mxGetN(prhs[0]);
    ssn_csr A; memset(&A, 0, sizeof(A)); ssn_mex_upload_sparse(c, prhs[0], &A);
    double *b = ssn_mex_to_device(c, prhs[1], N), *guess = NULL;
    ssn_amg_options o; ssn_mex_amg_options(c, prhs[2], N, &o, &guess);
    double *x = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    int it = 0, len = 0; double rel = 0;
    int st = ssn_class_amg(c, &A, b, &o, /*keep_hierarchy=*/1, x, &it, &rel, NULL, NULL, &len);
    plhs[0] = ssn_mex_from_device(c, x, N, &st);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(it);
    ssn_free(c, b); ssn_free(c, x); if (guess) ssn_free(c, guess); ssn_csr_free(c, &A);
    ssn_mex_check(st);
}

void *shim_ctx(void) { return (void *)g_ctx; }
long long shim_rng_drawn(void) { return (long long)ssn_rng_drawn(ssn_mex_ctx()); }
int shim_rng_reset(void) { return ssn_rng_reset(ssn_mex_ctx(), 5489u); }
int shim_clear(void) { return ssn_amg_clear(ssn_mex_ctx()); }
