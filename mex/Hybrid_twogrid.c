/* [zeta,itamg,resamg,info] = Hybrid_twogrid(prob_data,amg_options)
 * -- MEX replacement of the reference's Hybrid_twogrid.m:1-90 (inner_solver = 5). */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 2 || !mxIsStruct(prhs[0]) || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("ssnamg:nargin", "Hybrid_twogrid(prob_data,amg_options)");
    ssn_ctx *c = ssn_mex_ctx();
    ssn_mex_pd pd; ssn_mex_prob_data(c, prhs[0], 0, &pd);
    const size_t N = (size_t)(pd.d.m + pd.d.n);
    ssn_amg_options o; double *guess = NULL; ssn_mex_amg_options(c, prhs[1], 0, &o, &guess);   /* guess is overwritten inside, Hybrid_twogrid.m:38 */
    double *zeta = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    int it = 0, info[2] = {0, 0}; double res = 0;
    int st = ssn_hybrid_twogrid(c, &pd.d, &o, zeta, &it, &res, info);
    plhs[0] = ssn_mex_from_device(c, zeta, N, &st);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(it);
    if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(res);
    if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix(1, 2, mxREAL); mxGetPr(plhs[3])[0] = info[0]; mxGetPr(plhs[3])[1] = info[1]; }
    ssn_free(c, zeta); ssn_mex_prob_data_free(c, &pd);
    ssn_mex_check(st);                                          /* error raised after temporaries are gone */
}
