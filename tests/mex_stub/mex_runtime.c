/* mex_runtime.c -- a small FUNCTIONAL stand-in for the part of MATLAB's MEX C API that the shims under mex/ use
 * (declared in tests/mex_stub/mex.h), so that the real shim sources can be compiled into shared objects and
 * CALLED from a C host the way MATLAB calls .mex files (tests/c/mex_host.c).  Test infrastructure only:
 * full double / logical / sparse double arrays, double scalars, one-element structs, char scalars.
 * mexErrMsgIdAndTxt prints and longjmps back into the host's call wrapper (mex_call), like MATLAB unwinds. */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { K_DOUBLE, K_LOGICAL, K_SPARSE, K_STRUCT, K_CHAR };
struct mxArray_tag {
    int kind; size_t m, n, nzmax;
    double *pr; mxLogical *lg; mwIndex *jc, *ir; mxChar *ch;
    int nfields; const char *names[32]; mxArray *fields[32];
};

static jmp_buf *g_jmp = NULL;
char mex_last_error_id[128];
char mex_last_error_msg[1024];
static void (*g_atexit[64])(void);
static int g_natexit = 0, g_locks = 0;

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    snprintf(mex_last_error_id, sizeof(mex_last_error_id), "%s", id);
    vsnprintf(mex_last_error_msg, sizeof(mex_last_error_msg), fmt, ap);
    va_end(ap);
    if (g_jmp) longjmp(*g_jmp, 1);
    fprintf(stderr, "%s: %s\n", mex_last_error_id, mex_last_error_msg);
    exit(3);
}
void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...) { (void)id; (void)fmt; }
int mexAtExit(void (*fn)(void)) { if (g_natexit < 64) g_atexit[g_natexit++] = fn; return 0; }
void mexLock(void) { ++g_locks; }
int mex_runtime_locks(void) { return g_locks; }
void mex_runtime_run_atexit(void) { for (int i = g_natexit - 1; i >= 0; --i) g_atexit[i](); g_natexit = 0; }
/* returns 0, or 1 when the shim raised a MATLAB error (mex_last_error_*) */
int mex_call(void (*fn)(int, mxArray **, int, const mxArray **), int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs) {
    jmp_buf jb; g_jmp = &jb;
    if (setjmp(jb)) { g_jmp = NULL; return 1; }
    fn(nlhs, plhs, nrhs, prhs);
    g_jmp = NULL;
    return 0;
}

static mxArray *new_array(int kind, size_t m, size_t n) {
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray)); a->kind = kind; a->m = m; a->n = n; return a;
}
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
bool mxIsDouble(const mxArray *a) { return a->kind == K_DOUBLE || a->kind == K_SPARSE; }
bool mxIsSparse(const mxArray *a) { return a->kind == K_SPARSE; }
bool mxIsStruct(const mxArray *a) { return a->kind == K_STRUCT; }
bool mxIsChar(const mxArray *a) { return a->kind == K_CHAR; }
bool mxIsEmpty(const mxArray *a) { return a->m * a->n == 0; }
bool mxIsLogical(const mxArray *a) { return a->kind == K_LOGICAL; }
double *mxGetPr(const mxArray *a) { return a->pr; }
void *mxGetData(const mxArray *a) { return a->kind == K_CHAR ? (void *)a->ch : (a->kind == K_LOGICAL ? (void *)a->lg : (void *)a->pr); }
mxLogical *mxGetLogicals(const mxArray *a) { return a->lg; }
double mxGetScalar(const mxArray *a) { return a->kind == K_LOGICAL ? (double)a->lg[0] : (a->kind == K_CHAR ? (double)a->ch[0] : a->pr[0]); }
int mxGetString(const mxArray *a, char *buf, mwSize buflen) {
    if (a->kind != K_CHAR || buflen == 0) return 1;
    size_t k = a->m * a->n, i;
    if (k > buflen - 1) k = buflen - 1;
    for (i = 0; i < k; ++i) buf[i] = (char)a->ch[i];
    buf[k] = 0;
    return 0;
}
mwIndex *mxGetJc(const mxArray *a) { return a->jc; }
mwIndex *mxGetIr(const mxArray *a) { return a->ir; }
mwSize mxGetNzmax(const mxArray *a) { return a->nzmax; }
mxArray *mxGetField(const mxArray *s, mwIndex i, const char *name) {
    (void)i;
    for (int k = 0; k < s->nfields; ++k) if (!strcmp(s->names[k], name)) return s->fields[k];
    return NULL;
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    (void)c; mxArray *a = new_array(K_DOUBLE, m, n); a->pr = (double *)calloc((m * n) != 0 ? m * n : 1, sizeof(double)); return a;
}
mxArray *mxCreateDoubleScalar(double v) { mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL); a->pr[0] = v; return a; }
mxArray *mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity c) {
    (void)c; mxArray *a = new_array(K_SPARSE, m, n); a->nzmax = nzmax ? nzmax : 1;
    a->pr = (double *)calloc(a->nzmax, sizeof(double)); a->ir = (mwIndex *)calloc(a->nzmax, sizeof(mwIndex));
    a->jc = (mwIndex *)calloc(n + 1, sizeof(mwIndex)); return a;
}
mxArray *mxCreateLogicalMatrix(mwSize m, mwSize n) {
    mxArray *a = new_array(K_LOGICAL, m, n); a->lg = (mxLogical *)calloc((m * n) != 0 ? m * n : 1, sizeof(mxLogical)); return a;
}
/* host-side helpers that MATLAB itself would provide */
mxArray *mex_runtime_struct(void) { return new_array(K_STRUCT, 1, 1); }
void mex_runtime_set_field(mxArray *s, const char *name, mxArray *v) { s->names[s->nfields] = name; s->fields[s->nfields++] = v; }
mxArray *mex_runtime_char(char ch) { mxArray *a = new_array(K_CHAR, 1, 1); a->ch = (mxChar *)calloc(1, sizeof(mxChar)); a->ch[0] = (mxChar)ch; return a; }
void *mxMalloc(size_t n) { return malloc(n ? n : 1); }
void *mxCalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz ? sz : 1); }
void mxFree(void *p) { free(p); }
int mexCallMATLAB(int nlhs, mxArray *plhs[], int nrhs, mxArray *prhs[], const char *name) {
    (void)nlhs; (void)plhs; (void)nrhs; (void)prhs; (void)name; return 1;
}
void mxDestroyArray(mxArray *a) {
    if (!a) return;
    free(a->pr); free(a->lg); free(a->jc); free(a->ir); free(a->ch);
    for (int k = 0; k < a->nfields; ++k) mxDestroyArray(a->fields[k]);
    free(a);
}
const mxArray *mexGetVariablePtr(const char *workspace, const char *name) { (void)workspace; (void)name; return NULL; }
