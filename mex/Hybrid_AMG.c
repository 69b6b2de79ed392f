/* [zeta,itamg,resamg,info] = Hybrid_AMG(prob_data,amg_options)
 * -- MEX replacement of the reference's Hybrid_AMG.m:1-114. */
#include "ssn_mex_common.h"

static double fld(const mxArray *s, const char *name, double dflt) {
    const mxArray *f = mxGetField(s, 0, name);
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;      /* isempty() -> reference default */
}

/* MATLAB sparse (CSC, symmetric pattern) -> device CSR */
static void upload_sparse(ssn_ctx *c, const mxArray *S, ssn_csr *out) {
    const mwSize N = mxGetN(S); const mwIndex *jc = mxGetJc(S), *ir = mxGetIr(S);
    const mwIndex nnz = jc[N];
    int32_t *rp = (int32_t *)mxMalloc(sizeof(int32_t) * (N + 1)), *ci = (int32_t *)mxMalloc(sizeof(int32_t) * (nnz ? nnz : 1));
    for (mwSize k = 0; k <= N; ++k) rp[k] = (int32_t)jc[k];
    for (mwIndex k = 0; k < nnz; ++k) ci[k] = (int32_t)ir[k];
    int st = ssn_csr_upload(c, (int64_t)N, (int64_t)N, (int64_t)nnz, rp, ci, mxGetPr(S), out);
    mxFree(rp); mxFree(ci);
    ssn_mex_check(st);
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 2 || !mxIsStruct(prhs[0]) || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("ssnamg:nargin", "Hybrid_AMG(prob_data,amg_options)");
    ssn_ctx *c = ssn_mex_ctx();
    const mxArray *pd = prhs[0], *op = prhs[1];
    const mxArray *p = mxGetField(pd, 0, "p"), *q = mxGetField(pd, 0, "q"), *z = mxGetField(pd, 0, "z");
    const mxArray *H0 = mxGetField(pd, 0, "H0"), *T = mxGetField(pd, 0, "T");
    const size_t m = mxGetNumberOfElements(p), n = mxGetNumberOfElements(q), N = m + n;
    ssn_prob_data d; memset(&d, 0, sizeof(d));
    d.bk1 = mxGetScalar(mxGetField(pd, 0, "bk1")); d.tk = mxGetScalar(mxGetField(pd, 0, "tk"));
    d.m = (int64_t)m; d.n = (int64_t)n;
    double *pdv = ssn_mex_to_device(c, p, m), *qdv = ssn_mex_to_device(c, q, n), *zdv = ssn_mex_to_device(c, z, N);
    double *tdv = NULL;
    if (T && mxGetNzmax(T) > 0 && mxGetJc(T)[N] > 0) {          /* T = spdiags(t): pull the diagonal out */
        double *t = (double *)mxCalloc(N, sizeof(double));
        const mwIndex *jc = mxGetJc(T), *ir = mxGetIr(T); const double *pr = mxGetPr(T);
        for (size_t j = 0; j < N; ++j) for (mwIndex k = jc[j]; k < jc[j + 1]; ++k) if (ir[k] == j) t[j] = pr[k];
        void *dv = NULL; ssn_mex_check(ssn_malloc(c, N * sizeof(double), &dv));
        ssn_mex_check(ssn_memcpy_h2d(c, dv, t, N * sizeof(double))); tdv = (double *)dv; mxFree(t);
    }
    ssn_csr H; memset(&H, 0, sizeof(H)); upload_sparse(c, H0, &H);
    d.p_dev = pdv; d.q_dev = qdv; d.z_dev = zdv; d.t_dev = tdv; d.H0 = &H;
    ssn_amg_options o;
    o.retol = fld(op, "retol", -1); o.bigph = (int)fld(op, "bigph", -1); o.maxit = (int)fld(op, "maxit", -1);
    o.theta = fld(op, "theta", -1); o.smoth = (int)fld(op, "smoth", -1); o.isnsp = (int)fld(op, "isnsp", -1);
    o.inter = (int)fld(op, "inter", -1); o.fnode = 0; o.guess_dev = NULL;
    { const mxArray *cy = mxGetField(op, 0, "cycle");
      o.cycle = (cy && mxIsChar(cy)) ? (int)*(mxChar *)mxGetData(cy) : (cy && !mxIsEmpty(cy) ? (int)mxGetScalar(cy) : -1); }
    void *zeta = NULL; ssn_mex_check(ssn_malloc(c, N * sizeof(double), &zeta));
    int it = 0, info[2] = {0, 0}; double res = 0;
    int st = ssn_hybrid_amg(c, &d, &o, (double *)zeta, &it, &res, info);
    if (st == SSN_OK) {
        plhs[0] = mxCreateDoubleMatrix((mwSize)N, 1, mxREAL);
        st = ssn_memcpy_d2h(c, mxGetPr(plhs[0]), zeta, N * sizeof(double));
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(it);
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(res);
        if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix(1, 2, mxREAL); mxGetPr(plhs[3])[0] = info[0]; mxGetPr(plhs[3])[1] = info[1]; }
    }
    ssn_free(c, zeta); ssn_free(c, pdv); ssn_free(c, qdv); ssn_free(c, zdv); if (tdv) ssn_free(c, tdv); ssn_csr_free(c, &H);
    ssn_mex_check(st);                                          /* error raised after temporaries are gone */
}
