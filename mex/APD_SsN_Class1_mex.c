/* [xk,lk,fxk,KKT_xk,KKT_lk,info] = APD_SsN_Class1_mex(c,r,l,p,q,gama,inner_solver,maxit,KKT_Tol) -- the body of the
 * reference's SCRIPT Class1/APD_SsN_Class1.m:32-275 (warm start :59, APD outer loop :101-275, SsN inner loop :137-238,
 * line search :182-211, KKT bookkeeping :239-274) as ONE MEX call: the host arrays go to the device once, the plan and
 * the duals come back once (ssn_apd_ssn_class1_host).  A thin `APD_SsN_Class1.m` wrapper that loads the data (:27),
 * calls this and draws the figures (:276-331) keeps the demo's behaviour.  gama: scalar (Inf) or an m*n vector;
 * inner_solver 2 / 3 / 4 (default) / 5 as at :66-70.  info = [outer_its converged rel_kkt ssn_steps ls_trials amg_calls
 * warmup_s loop_s]. */
#include <math.h>
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 5) mexErrMsgIdAndTxt("ssnamg:nargin", "[xk,lk,fxk,KKT_xk,KKT_lk,info] = APD_SsN_Class1_mex(c,r,l,p,q,gama,inner_solver,maxit,KKT_Tol)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[3]), n = mxGetNumberOfElements(prhs[4]);
    if (mxGetNumberOfElements(prhs[0]) != m * n || mxGetNumberOfElements(prhs[1]) != n || mxGetNumberOfElements(prhs[2]) != m)
        mexErrMsgIdAndTxt("ssnamg:arg", "c must have m*n, r n and l m elements");
    const int gama_vec = nrhs > 5 && mxGetNumberOfElements(prhs[5]) == m * n && m * n > 1;
    const double gama_s = (nrhs > 5 && !gama_vec && !mxIsEmpty(prhs[5])) ? mxGetScalar(prhs[5]) : INFINITY;
    ssn_apd_options o; memset(&o, 0, sizeof(o));
    o.inner_solver = nrhs > 6 ? (int)mxGetScalar(prhs[6]) : 4;
    o.maxit = nrhs > 7 ? (int)mxGetScalar(prhs[7]) : 100;
    o.KKT_Tol = nrhs > 8 ? mxGetScalar(prhs[8]) : 1e-6;
    o.warm_maxit = -1;
    const size_t hist = (size_t)o.maxit + 1;
    plhs[0] = mxCreateDoubleMatrix((mwSize)(m * n), 1, mxREAL);
    mxArray *lk = mxCreateDoubleMatrix((mwSize)(m + n), 1, mxREAL);
    double *fx = (double *)mxCalloc(hist, sizeof(double)), *kx = (double *)mxCalloc(hist, sizeof(double)), *kl = (double *)mxCalloc(hist, sizeof(double));
    ssn_apd_result res; memset(&res, 0, sizeof(res));
    const int st = ssn_apd_ssn_class1_host(c, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]),
                                           (int64_t)m, (int64_t)n, gama_vec ? mxGetPr(prhs[5]) : NULL, gama_s, &o, mxGetPr(plhs[0]), mxGetPr(lk),
                                           &res, fx, kx, kl, NULL, NULL, 0);
    if (nlhs > 1) plhs[1] = lk;
    const size_t L = st == SSN_OK ? (size_t)res.hist_len : 0;
    if (nlhs > 2) { plhs[2] = mxCreateDoubleMatrix((mwSize)L, 1, mxREAL); memcpy(mxGetPr(plhs[2]), fx, sizeof(double) * L); }
    if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix((mwSize)L, 1, mxREAL); memcpy(mxGetPr(plhs[3]), kx, sizeof(double) * L); }
    if (nlhs > 4) { plhs[4] = mxCreateDoubleMatrix((mwSize)L, 1, mxREAL); memcpy(mxGetPr(plhs[4]), kl, sizeof(double) * L); }
    if (nlhs > 5) {
        plhs[5] = mxCreateDoubleMatrix(1, 8, mxREAL);
        double *v = mxGetPr(plhs[5]);
        v[0] = res.outer_its; v[1] = res.converged; v[2] = res.rel_kkt; v[3] = res.ssn_steps; v[4] = res.ls_trials; v[5] = res.amg_calls;
        v[6] = res.warmup_s; v[7] = res.loop_s;
    }
    mxFree(fx); mxFree(kx); mxFree(kl);
    ssn_mex_check(st);
}
