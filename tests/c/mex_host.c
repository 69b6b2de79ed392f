/* mex_host.c -- calls the MEX shims the way MATLAB does: every shim is its own shared object with its own copy of
 * ssn_mex_common.h's static state, dlopen'ed separately (RTLD_LOCAL), all linked against the one libssnamg.so.
 * Checks that they share ONE context (hierarchy handle + MT19937 stream), like the reference's functions share
 * `global Ack Prok J smoth_it Rk` (AMG/Class_AMG.m:43, AMG/MG_Wcycle.m:9) and MATLAB's global rand stream
 * (AMG/mis_set.m:35):
 *   1. keep.mex   : Class_AMG setup + solve with the hierarchy kept
 *   2. MG_Wcycle.mex (the real mex/MG_Wcycle.c): one W-cycle on THAT hierarchy through its own handle
 *   3. mis_set.mex (the real mex/mis_set.c): draws N numbers from the stream keep.mex sees
 *   4. clearing the hierarchy through one shim makes MG_Wcycle.mex raise the reference's error condition
 * usage: mex_host <dir with the shim .so files>;  exit 0 = all checks passed, 3 = no CUDA device. */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"
#include "../../include/ssnamg.h"

typedef void (*mexfn)(int, mxArray **, int, const mxArray **);
int mex_call(mexfn fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs);
void mex_runtime_run_atexit(void);
int mex_runtime_locks(void);
mxArray *mex_runtime_struct(void);
void mex_runtime_set_field(mxArray *s, const char *name, mxArray *v);
mxArray *mex_runtime_char(char ch);
extern char mex_last_error_id[128];
extern char mex_last_error_msg[1024];

static void *load(const char *dir, const char *name) {
    char path[4096]; snprintf(path, sizeof(path), "%s/%s", dir, name);
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen %s: %s\n", path, dlerror()); exit(2); }
    return h;
}
#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAILED: " __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)

int main(int argc, char **argv) {
    const char *dir = argc > 1 ? argv[1] : ".";
    void *hk = load(dir, "keep.mex.so"), *hw = load(dir, "MG_Wcycle.mex.so"), *hm = load(dir, "mis_set.mex.so");
    mexfn keep = (mexfn)dlsym(hk, "mexFunction"), wcyc = (mexfn)dlsym(hw, "mexFunction"), mis = (mexfn)dlsym(hm, "mexFunction");
    void *(*keep_ctx)(void) = (void *(*)(void))dlsym(hk, "shim_ctx");
    long long (*keep_drawn)(void) = (long long (*)(void))dlsym(hk, "shim_rng_drawn");
    int (*keep_reset)(void) = (int (*)(void))dlsym(hk, "shim_rng_reset");
    int (*keep_clear)(void) = (int (*)(void))dlsym(hk, "shim_clear");
    CHECK(keep && wcyc && mis && keep_ctx && keep_drawn && keep_reset && keep_clear, "missing symbols");
    CHECK(keep != wcyc && wcyc != mis, "the shims must be distinct objects");

    /* A = 5-point Laplacian on a g x g grid + 0.05*I (symmetric: CSC == CSR), b = ones */
    const int g = 24, N = g * g;
    mxArray *A = mxCreateSparse(N, N, 5 * N, mxREAL), *b = mxCreateDoubleMatrix(N, 1, mxREAL);
    { mwIndex *jc = mxGetJc(A), *ir = mxGetIr(A); double *pr = mxGetPr(A); size_t k = 0;
      for (int j = 0; j < N; ++j) {
          const int a = j / g, c = j % g; jc[j] = k;
          if (a > 0) { ir[k] = j - g; pr[k++] = -1; }
          if (c > 0) { ir[k] = j - 1; pr[k++] = -1; }
          ir[k] = j; pr[k++] = 4.05;
          if (c < g - 1) { ir[k] = j + 1; pr[k++] = -1; }
          if (a < g - 1) { ir[k] = j + g; pr[k++] = -1; }
      }
      jc[N] = k; }
    for (int i = 0; i < N; ++i) mxGetPr(b)[i] = 1.0;
    mxArray *o = mex_runtime_struct();
    mex_runtime_set_field(o, "retol", mxCreateDoubleScalar(1e-10)); mex_runtime_set_field(o, "bigph", mxCreateDoubleScalar(0));
    mex_runtime_set_field(o, "maxit", mxCreateDoubleScalar(30)); mex_runtime_set_field(o, "theta", mxCreateDoubleScalar(0.25));
    mex_runtime_set_field(o, "smoth", mxCreateDoubleScalar(2)); mex_runtime_set_field(o, "cycle", mex_runtime_char('w'));
    mex_runtime_set_field(o, "isnsp", mxCreateDoubleScalar(0)); mex_runtime_set_field(o, "inter", mxCreateDoubleScalar(1));

    /* 1. keep.mex */
    mxArray *out[3] = {0, 0, 0}; const mxArray *in3[3] = {A, b, o};
    if (mex_call(keep, 2, out, 3, in3)) {
        if (!strcmp(mex_last_error_id, "ssnamg:nogpu")) { fprintf(stderr, "no CUDA device: %s\n", mex_last_error_msg); return 3; }
        CHECK(0, "keep.mex: %s", mex_last_error_msg);
    }
    const int its = (int)mxGetScalar(out[1]);
    CHECK(its >= 2 && its <= 30, "Class_AMG cycles = %d", its);      /* the reference's AMG is slow on this matrix (0.86 per W-cycle): 30 = maxit */
    CHECK(ssn_default_ctx_refcount() == 1, "refcount after the first shim = %d", ssn_default_ctx_refcount());

    /* 2. the real MG_Wcycle.mex on the hierarchy keep.mex left behind: e = MG_Wcycle(b, 0, 1) */
    mxArray *e[1] = {0}; mxArray *isnsp = mxCreateDoubleScalar(0), *kk = mxCreateDoubleScalar(1);
    const mxArray *inw[3] = {b, isnsp, kk};
    CHECK(mex_call(wcyc, 1, e, 3, inw) == 0, "MG_Wcycle.mex through its own handle: %s (separate contexts?)", mex_last_error_msg);
    CHECK(ssn_default_ctx_refcount() == 2, "refcount after the second shim = %d", ssn_default_ctx_refcount());
    { /* a W-cycle from a zero guess reduces the residual of A e = b (by the factor the oracle's first cycle shows) */
      const mwIndex *jc = mxGetJc(A), *ir = mxGetIr(A); const double *pr = mxGetPr(A), *ev = mxGetPr(e[0]);
      double r2 = 0, b2 = 0;
      for (int i = 0; i < N; ++i) { double s = 0; for (mwIndex k = jc[i]; k < jc[i + 1]; ++k) s += pr[k] * ev[ir[k]]; r2 += (1.0 - s) * (1.0 - s); b2 += 1.0; }
      CHECK(sqrt(r2 / b2) < 0.95 && sqrt(r2 / b2) > 0.5, "relative residual after one W-cycle: %g (oracle: 0.86)", sqrt(r2 / b2)); }

    /* 3. one random stream: mis_set.mex draws N numbers (all nodes are connected) from the stream keep.mex sees */
    CHECK(keep_reset() == SSN_OK, "rng reset");
    const long long d0 = keep_drawn();
    mxArray *cf[2] = {0, 0}; mxArray *theta = mxCreateDoubleScalar(0.25); const mxArray *inm[2] = {A, theta};
    CHECK(mex_call(mis, 2, cf, 2, inm) == 0, "mis_set.mex: %s", mex_last_error_msg);
    const long long d1 = keep_drawn();
    CHECK(d0 == 0 && d1 - d0 == N, "draws seen through the other shim: %lld -> %lld (expected +%d)", d0, d1, N);
    CHECK(ssn_default_ctx_refcount() == 3, "refcount after the third shim = %d", ssn_default_ctx_refcount());
    { int nc = 0, nf = 0; for (int i = 0; i < N; ++i) { nc += mxGetLogicals(cf[0])[i]; nf += mxGetLogicals(cf[1])[i]; }
      CHECK(nc > 0 && nf > 0 && nc + nf == N, "C/F split: %d + %d != %d", nc, nf, N); }

    /* 4. clearing the hierarchy through one shim is seen by the other (Class_AMG.m:110 `clear global`) */
    CHECK(keep_clear() == SSN_OK, "clear");
    mxArray *e2[1] = {0};
    CHECK(mex_call(wcyc, 1, e2, 3, inw) == 1, "MG_Wcycle.mex after the clear must raise an error");
    CHECK(strstr(mex_last_error_msg, "hierarchy") != NULL, "unexpected error text: %s", mex_last_error_msg);

    CHECK(mex_runtime_locks() == 3, "mexLock calls = %d", mex_runtime_locks());
    mex_runtime_run_atexit();                              /* MATLAB exit: every shim drops its reference */
    CHECK(ssn_default_ctx_refcount() == 0, "refcount at exit = %d", ssn_default_ctx_refcount());
    printf("OK: %d W-cycles, 3 shims, one context %p, %lld shared draws\n", its, keep_ctx(), d1 - d0);
    return 0;
}
