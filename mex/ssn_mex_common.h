/* ssn_mex_common.h -- shared helpers of the MEX shims (compile only where MATLAB's mex.h exists;
 * this image has neither MATLAB nor Octave, so the shims are provided as source, type-checked
 * against a declaration-only mex.h stub by tests/test_abi.py, and the ctypes mirror api.py is the
 * executable stand-in).  One shim per reference function name: a `Name.mex*` on the MATLAB path
 * shadows `Name.m`, which is the drop-in mechanism. */
#ifndef SSN_MEX_COMMON_H
#define SSN_MEX_COMMON_H
#include <string.h>
#include "mex.h"
#include "matrix.h"
#include "../include/ssnamg.h"

/* This MEX file's reference to the PROCESS-WIDE default context of libssnamg.so.  Every shim (Class_AMG.mex,
 * MG_Wcycle.mex, transfer.mex, mis_set.mex, Hybrid_AMG.mex, ...) holds the same handle, so they share one AMG
 * hierarchy and one MT19937 stream exactly like the reference's functions share `global Ack Prok J smoth_it Rk`
 * (AMG/Class_AMG.m:43,110; AMG/MG_Wcycle.m:9; AMG/transfer.m:17) and MATLAB's global `rand`. */
static ssn_ctx *g_ctx = NULL;

static void ssn_mex_cleanup(void) { if (g_ctx) { ssn_default_ctx_release(); g_ctx = NULL; } }

/* keep the MEX file locked while it holds a reference, drop it at exit (the last shim to exit frees the context) */
static ssn_ctx *ssn_mex_ctx(void) {
    if (!g_ctx) {
        if (ssn_default_ctx_acquire(&g_ctx) != SSN_OK) mexErrMsgIdAndTxt("ssnamg:nogpu", "no CUDA device (no CPU fallback)");
        mexAtExit(ssn_mex_cleanup);
        mexLock();
    }
    return g_ctx;
}

/* status -> MATLAB error, raised only after temporaries were released by the caller */
static void ssn_mex_check(int st) {
    if (st != SSN_OK) mexErrMsgIdAndTxt("ssnamg:error", "%s", ssn_last_error(g_ctx));
}

static void *ssn_mex_dev_alloc(ssn_ctx *c, size_t bytes) {
    void *d = NULL;
    ssn_mex_check(ssn_malloc(c, bytes ? bytes : 8, &d));
    return d;
}

static double *ssn_mex_to_device(ssn_ctx *c, const mxArray *a, size_t count) {
    void *d = NULL;
    if (mxGetNumberOfElements(a) != count || !mxIsDouble(a) || mxIsSparse(a))
        mexErrMsgIdAndTxt("ssnamg:arg", "expected a full double array of %zu elements", count);
    ssn_mex_check(ssn_malloc(c, count * sizeof(double), &d));
    ssn_mex_check(ssn_memcpy_h2d(c, d, mxGetPr(a), count * sizeof(double)));
    return (double *)d;
}

/* device vector -> MATLAB column vector */
static mxArray *ssn_mex_from_device(ssn_ctx *c, const double *d, size_t count, int *st) {
    mxArray *a = mxCreateDoubleMatrix((mwSize)count, 1, mxREAL);
    if (*st == SSN_OK) *st = ssn_memcpy_d2h(c, mxGetPr(a), d, count * sizeof(double));
    return a;
}

/* device uint8 flags -> MATLAB logical column vector */
static mxArray *ssn_mex_logical_from_device(ssn_ctx *c, const uint8_t *d, size_t count, int *st) {
    mxArray *a = mxCreateLogicalMatrix((mwSize)count, 1);
    if (*st == SSN_OK) *st = ssn_memcpy_d2h(c, mxGetLogicals(a), d, count);
    return a;
}

/* device CSR (symmetric pattern, or the transpose is wanted) -> MATLAB sparse (CSC, mwIndex) */
static mxArray *ssn_mex_csr_to_sparse(ssn_ctx *c, const ssn_csr *A) {
    mxArray *S = mxCreateSparse((mwSize)A->ncols, (mwSize)A->nrows, (mwSize)(A->nnz ? A->nnz : 1), mxREAL);
    int32_t *rp = (int32_t *)mxMalloc(sizeof(int32_t) * (A->nrows + 1));
    int32_t *ci = (int32_t *)mxMalloc(sizeof(int32_t) * (A->nnz ? A->nnz : 1));
    ssn_mex_check(ssn_csr_download(c, A, rp, ci, mxGetPr(S)));
    mwIndex *jc = mxGetJc(S), *ir = mxGetIr(S);
    for (int64_t k = 0; k <= A->nrows; ++k) jc[k] = (mwIndex)rp[k];
    for (int64_t k = 0; k < A->nnz; ++k) ir[k] = (mwIndex)ci[k];
    mxFree(rp); mxFree(ci);
    return S;
}

/* a general (non-symmetric) device CSR -> MATLAB sparse: transpose on the device first, so the
 * CSR arrays of A' are the CSC arrays of A */
static mxArray *ssn_mex_csr_to_sparse_general(ssn_ctx *c, const ssn_csr *A, int *st) {
    ssn_csr At; memset(&At, 0, sizeof(At));
    mxArray *S = NULL;
    if (*st == SSN_OK) *st = ssn_transpose(c, A, &At);
    if (*st == SSN_OK) S = ssn_mex_csr_to_sparse(c, &At);
    ssn_csr_free(c, &At);
    return S;
}

/* MATLAB sparse (CSC) of a structurally symmetric matrix -> device CSR (same arrays) */
static void ssn_mex_upload_sparse(ssn_ctx *c, const mxArray *S, ssn_csr *out) {
    const mwSize N = mxGetN(S); const mwIndex *jc = mxGetJc(S), *ir = mxGetIr(S);
    const mwIndex nnz = jc[N];
    int32_t *rp = (int32_t *)mxMalloc(sizeof(int32_t) * (N + 1)), *ci = (int32_t *)mxMalloc(sizeof(int32_t) * (nnz ? nnz : 1));
    for (mwSize k = 0; k <= N; ++k) rp[k] = (int32_t)jc[k];
    for (mwIndex k = 0; k < nnz; ++k) ci[k] = (int32_t)ir[k];
    int st = ssn_csr_upload(c, (int64_t)N, (int64_t)mxGetM(S), (int64_t)nnz, rp, ci, mxGetPr(S), out);
    mxFree(rp); mxFree(ci);
    ssn_mex_check(st);
}

static double ssn_mex_fld(const mxArray *s, const char *name, double dflt) {
    const mxArray *f = s ? mxGetField(s, 0, name) : NULL;
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;      /* isempty() -> reference default */
}

/* amg_options struct (AMG/Class_AMG.m:20-34) -> ssn_amg_options; *guess_dev receives an owned device copy */
static void ssn_mex_amg_options(ssn_ctx *c, const mxArray *op, size_t N, ssn_amg_options *o, double **guess_dev) {
    o->retol = ssn_mex_fld(op, "retol", -1); o->bigph = (int)ssn_mex_fld(op, "bigph", -1); o->maxit = (int)ssn_mex_fld(op, "maxit", -1);
    o->theta = ssn_mex_fld(op, "theta", -1); o->smoth = (int)ssn_mex_fld(op, "smoth", -1); o->isnsp = (int)ssn_mex_fld(op, "isnsp", -1);
    o->inter = (int)ssn_mex_fld(op, "inter", -1); o->fnode = (int)ssn_mex_fld(op, "fnode", 0); o->guess_dev = NULL;
    { const mxArray *cy = op ? mxGetField(op, 0, "cycle") : NULL;
      o->cycle = (cy && mxIsChar(cy)) ? (int)*(mxChar *)mxGetData(cy) : (cy && !mxIsEmpty(cy) ? (int)mxGetScalar(cy) : -1); }
    *guess_dev = NULL;
    { const mxArray *g = op ? mxGetField(op, 0, "guess") : NULL;
      if (g && !mxIsEmpty(g) && N) { *guess_dev = ssn_mex_to_device(c, g, N); o->guess_dev = *guess_dev; } }
}

/* pcg_options struct (PCG.m:18-27) */
static void ssn_mex_pcg_options(ssn_ctx *c, const mxArray *op, size_t N, ssn_pcg_options *o, double **guess_dev) {
    o->retol = ssn_mex_fld(op, "retol", -1); o->maxit = (int)ssn_mex_fld(op, "maxit", -1);
    o->precd = (int)ssn_mex_fld(op, "precd", -1); o->nf = (int)ssn_mex_fld(op, "nf", 0); o->guess_dev = NULL;
    *guess_dev = NULL;
    { const mxArray *g = op ? mxGetField(op, 0, "guess") : NULL;
      if (g && !mxIsEmpty(g) && N) { *guess_dev = ssn_mex_to_device(c, g, N); o->guess_dev = *guess_dev; } }
}

/* prob_data struct (Class1/APD_SsN_Class1.m:154-156; + s, phi in Class2 :163-166) on the device */
typedef struct { ssn_prob_data d; ssn_csr H; double *p, *q, *z, *t, *phi; uint8_t *s; } ssn_mex_pd;

static void ssn_mex_prob_data(ssn_ctx *c, const mxArray *pd, int pot, ssn_mex_pd *out) {
    memset(out, 0, sizeof(*out));
    const mxArray *p = mxGetField(pd, 0, "p"), *q = mxGetField(pd, 0, "q"), *z = mxGetField(pd, 0, "z");
    const mxArray *H0 = mxGetField(pd, 0, "H0"), *T = mxGetField(pd, 0, "T");
    const size_t m = mxGetNumberOfElements(p), n = mxGetNumberOfElements(q), N = m + n;
    out->d.bk1 = mxGetScalar(mxGetField(pd, 0, "bk1")); out->d.tk = mxGetScalar(mxGetField(pd, 0, "tk"));
    out->d.m = (int64_t)m; out->d.n = (int64_t)n;
    out->p = ssn_mex_to_device(c, p, m); out->q = ssn_mex_to_device(c, q, n); out->z = ssn_mex_to_device(c, z, N + (pot ? 1 : 0));
    if (T && mxGetNzmax(T) > 0 && mxGetJc(T)[N] > 0) {          /* T = spdiags(t): pull the diagonal out */
        double *t = (double *)mxCalloc(N, sizeof(double));
        const mwIndex *jc = mxGetJc(T), *ir = mxGetIr(T); const double *pr = mxGetPr(T);
        for (size_t j = 0; j < N; ++j) for (mwIndex k = jc[j]; k < jc[j + 1]; ++k) if (ir[k] == j) t[j] = pr[k];
        out->t = (double *)ssn_mex_dev_alloc(c, N * sizeof(double));
        ssn_mex_check(ssn_memcpy_h2d(c, out->t, t, N * sizeof(double))); mxFree(t);
    }
    ssn_mex_upload_sparse(c, H0, &out->H);
    if (pot) {
        const mxArray *s = mxGetField(pd, 0, "s"), *phi = mxGetField(pd, 0, "phi");
        out->s = (uint8_t *)ssn_mex_dev_alloc(c, m * n);
        ssn_mex_check(ssn_memcpy_h2d(c, out->s, mxGetLogicals(s), m * n));
        out->phi = ssn_mex_to_device(c, phi, m * n);
    }
    out->d.p_dev = out->p; out->d.q_dev = out->q; out->d.z_dev = out->z; out->d.t_dev = out->t; out->d.H0 = &out->H;
    out->d.s_dev = out->s; out->d.phi_dev = out->phi;
}
static void ssn_mex_prob_data_free(ssn_ctx *c, ssn_mex_pd *x) {
    ssn_free(c, x->p); ssn_free(c, x->q); ssn_free(c, x->z);
    if (x->t) ssn_free(c, x->t);
    if (x->s) ssn_free(c, x->s);
    if (x->phi) ssn_free(c, x->phi);
    ssn_csr_free(c, &x->H);
}
#endif
