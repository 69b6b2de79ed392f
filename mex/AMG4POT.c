/* [zeta,it,res,info] = AMG4POT(prob_data,amg_options,str)  (str = 'amg': the two solves through Hybrid_AMG; any other
 * string, e.g. 'twogrid': through Hybrid_twogrid, :45-51) -- MEX replacement of the reference's Class2/AMG4POT.m:1-56. */
#include <string.h>
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 2 || !mxIsStruct(prhs[0]) || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("ssnamg:nargin", "AMG4POT(prob_data,options)");
    ssn_ctx *c = ssn_mex_ctx();
    ssn_mex_pd pd; ssn_mex_prob_data(c, prhs[0], 1, &pd);
    const size_t N = (size_t)(pd.d.m + pd.d.n) + 1;
    ssn_amg_options o; double *guess = NULL; ssn_mex_amg_options(c, prhs[1], 0, &o, &guess);
    double *zeta = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
    int it = 0, info[2] = {0, 0}; double res = 0;
    int twogrid = 0;
    if (nrhs > 2 && mxIsChar(prhs[2])) { char buf[16] = {0}; mxGetString(prhs[2], buf, sizeof(buf)); twogrid = strcmp(buf, "amg") != 0; }
    int st = ssn_amg4pot_str(c, &pd.d, &o, twogrid, zeta, &it, &res, info);
    plhs[0] = ssn_mex_from_device(c, zeta, N, &st);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(it);
    if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(res);
    if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix(1, 2, mxREAL); mxGetPr(plhs[3])[0] = info[0]; mxGetPr(plhs[3])[1] = info[1]; }
    ssn_free(c, zeta); if (guess) ssn_free(c, guess); ssn_mex_prob_data_free(c, &pd);
    ssn_mex_check(st);
}
