/* ssn_mex_common.h -- shared helpers of the MEX shims (compile only where MATLAB's mex.h exists;
 * this image has neither MATLAB nor Octave, so the shims are provided as source and the ctypes
 * mirror api.py is the executable stand-in).  One shim per reference function name: a
 * `Name.mex*` on the MATLAB path shadows `Name.m`, which is the drop-in mechanism
 * (SURVEY.md section 8b). */
#ifndef SSN_MEX_COMMON_H
#define SSN_MEX_COMMON_H
#include <string.h>
#include "mex.h"
#include "matrix.h"
#include "../include/ssnamg.h"

static ssn_ctx *g_ctx = NULL;

static void ssn_mex_cleanup(void) { if (g_ctx) { ssn_destroy(g_ctx); g_ctx = NULL; } }

/* library-owned persistent state (hierarchy handle, random stream) lives in the context:
 * keep the MEX file locked, free at exit (the reference keeps it in globals,
 * AMG/Class_AMG.m:42-43) */
static ssn_ctx *ssn_mex_ctx(void) {
    if (!g_ctx) {
        if (ssn_create(&g_ctx, -1) != SSN_OK) mexErrMsgIdAndTxt("ssnamg:nogpu", "no CUDA device (no CPU fallback)");
        mexAtExit(ssn_mex_cleanup);
        mexLock();
    }
    return g_ctx;
}

/* status -> MATLAB error, raised only after temporaries were released by the caller */
static void ssn_mex_check(int st) {
    if (st != SSN_OK) mexErrMsgIdAndTxt("ssnamg:error", "%s", ssn_last_error(g_ctx));
}

static double *ssn_mex_to_device(ssn_ctx *c, const mxArray *a, size_t count) {
    void *d = NULL;
    if (mxGetNumberOfElements(a) != count || !mxIsDouble(a) || mxIsSparse(a))
        mexErrMsgIdAndTxt("ssnamg:arg", "expected a full double array of %zu elements", count);
    ssn_mex_check(ssn_malloc(c, count * sizeof(double), &d));
    ssn_mex_check(ssn_memcpy_h2d(c, d, mxGetPr(a), count * sizeof(double)));
    return (double *)d;
}

/* device CSR (symmetric pattern) -> MATLAB sparse (CSC, mwIndex) */
static mxArray *ssn_mex_csr_to_sparse(ssn_ctx *c, const ssn_csr *A) {
    mxArray *S = mxCreateSparse((mwSize)A->nrows, (mwSize)A->ncols, (mwSize)(A->nnz ? A->nnz : 1), mxREAL);
    int32_t *rp = (int32_t *)mxMalloc(sizeof(int32_t) * (A->nrows + 1));
    int32_t *ci = (int32_t *)mxMalloc(sizeof(int32_t) * (A->nnz ? A->nnz : 1));
    ssn_mex_check(ssn_csr_download(c, A, rp, ci, mxGetPr(S)));
    mwIndex *jc = mxGetJc(S), *ir = mxGetIr(S);
    for (int64_t k = 0; k <= A->nrows; ++k) jc[k] = (mwIndex)rp[k];
    for (int64_t k = 0; k < A->nnz; ++k) ir[k] = (mwIndex)ci[k];
    mxFree(rp); mxFree(ci);
    return S;
}
#endif
