/* [uk,lk,fxk,KKT,info] = APD_SsN_Class2_mex(c,r,l,p,q,mu,phi,inner_solver,maxit,KKT_Tol) -- the body of the reference's
 * SCRIPT Class2/APD_SsN_Class2.m:25-285 (warm start :50, APD outer loop :95-285, SsN inner loop :136-229, line search
 * :196-213, KKT bookkeeping :231-274, restart rule :253-257) as ONE MEX call: the host arrays go to the device once,
 * uk = [x;y;z] and the duals come back once (ssn_apd_ssn_class2_host).  inner_solver 3 (PCG4POT) / 4 (AMG4POT, default) /
 * 5 (AMG4POT with 'twogrid') as at :66-70.  KKT is 4-by-(outer_its+1): rows KKT_xk, KKT_yk, KKT_zk, KKT_lk.
 * info = [outer_its converged rel_kkt ssn_steps ls_trials amg_calls warmup_s loop_s]. */
#include "ssn_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 7) mexErrMsgIdAndTxt("ssnamg:nargin", "[uk,lk,fxk,KKT,info] = APD_SsN_Class2_mex(c,r,l,p,q,mu,phi,inner_solver,maxit,KKT_Tol)");
    ssn_ctx *c = ssn_mex_ctx();
    const size_t m = mxGetNumberOfElements(prhs[3]), n = mxGetNumberOfElements(prhs[4]);
    if (mxGetNumberOfElements(prhs[0]) != m * n || mxGetNumberOfElements(prhs[1]) != n || mxGetNumberOfElements(prhs[2]) != m ||
        mxGetNumberOfElements(prhs[6]) != m * n)
        mexErrMsgIdAndTxt("ssnamg:arg", "c and phi must have m*n, r n and l m elements");
    ssn_apd_options o; memset(&o, 0, sizeof(o));
    o.inner_solver = nrhs > 7 ? (int)mxGetScalar(prhs[7]) : 4;
    o.maxit = nrhs > 8 ? (int)mxGetScalar(prhs[8]) : 100;
    o.KKT_Tol = nrhs > 9 ? mxGetScalar(prhs[9]) : 1e-6;
    o.warm_maxit = -1;
    const size_t hist = (size_t)o.maxit + 1;
    plhs[0] = mxCreateDoubleMatrix((mwSize)(m * n + n + m), 1, mxREAL);
    mxArray *lk = mxCreateDoubleMatrix((mwSize)(m + n + 1), 1, mxREAL);
    double *fx = (double *)mxCalloc(hist, sizeof(double)), *kk = (double *)mxCalloc(4 * hist, sizeof(double));
    ssn_apd_result res; memset(&res, 0, sizeof(res));
    const int st = ssn_apd_ssn_class2_host(c, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]),
                                           (int64_t)m, (int64_t)n, mxGetScalar(prhs[5]), mxGetPr(prhs[6]), &o, mxGetPr(plhs[0]), mxGetPr(lk),
                                           &res, fx, kk, NULL, NULL, 0);
    if (nlhs > 1) plhs[1] = lk;
    const size_t L = st == SSN_OK ? (size_t)res.hist_len : 0;
    if (nlhs > 2) { plhs[2] = mxCreateDoubleMatrix((mwSize)L, 1, mxREAL); memcpy(mxGetPr(plhs[2]), fx, sizeof(double) * L); }
    if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix(4, (mwSize)L, mxREAL); memcpy(mxGetPr(plhs[3]), kk, sizeof(double) * 4 * L); }
    if (nlhs > 4) {
        plhs[4] = mxCreateDoubleMatrix(1, 8, mxREAL);
        double *v = mxGetPr(plhs[4]);
        v[0] = res.outer_its; v[1] = res.converged; v[2] = res.rel_kkt; v[3] = res.ssn_steps; v[4] = res.ls_trials; v[5] = res.amg_calls;
        v[6] = res.warmup_s; v[7] = res.loop_s;
    }
    mxFree(fx); mxFree(kk);
    ssn_mex_check(st);
}
