/* warmstart_from_mat.c -- a host program without MATLAB or Python: reads one of the reference's problem files
 * (Class1/InputData/data1-500.mat) with libssnmat.so, runs the A-ADMM warm start of Class1/warmup_class1.m on
 * the GPU through libssnamg.so and reports the marginal residual ||Ax(xk) - [r;l]|| / ||[r;l]|| and c'xk.
 *
 *   gcc -O2 -Iinclude examples/warmstart_from_mat.c -o warmstart_from_mat \
 *       -Lcodes-of-ipd-ssn-amg-method_b200 -lssnamg -lssnmat -lm -Wl,-rpath,'$ORIGIN/codes-of-ipd-ssn-amg-method_b200'
 *   ./warmstart_from_mat data1-500.mat [iterations = 100]
 *
 * Exit status: 0 ok, 1 usage, 2 input file, 3 GPU library (no CUDA device, out of memory, ...).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ssnamg.h"
#include "ssnamg_io.h"

#define GPU(call) do { int st_ = (call); if (st_ != SSN_OK) { \
    fprintf(stderr, "%s: status %d: %s\n", #call, st_, ctx ? ssn_last_error(ctx) : "no context"); rc = 3; goto done; } } while (0)

static int all_inf(const double* g, int64_t k) {
    if (!g) return 1;
    for (int64_t i = 0; i < k; ++i) if (!(isinf(g[i]) && g[i] > 0)) return 0;
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s problem.mat [iterations]\n", argv[0]); return 1; }
    const int maxit = argc > 2 ? atoi(argv[2]) : 100;
    ssn_problem pb;
    int st = ssn_problem_load(argv[1], &pb);
    if (st != SSN_MAT_OK) { fprintf(stderr, "%s: %s\n", argv[1], ssn_mat_strerror(st)); return 2; }
    const int64_t m = pb.m, n = pb.n, mn = m * n, N = m + n;
    printf("%s: m = %lld, n = %lld, gama %s\n", argv[1], (long long)m, (long long)n, all_inf(pb.gama, mn) ? "= +Inf" : "finite");

    int rc = 0;
    ssn_ctx* ctx = NULL;
    void *c_d = NULL, *b_d = NULL, *p_d = NULL, *q_d = NULL, *g_d = NULL, *x_d = NULL, *lk_d = NULL, *ax_d = NULL;
    double *b = (double*)malloc(sizeof(double) * (size_t)N), *ax = (double*)malloc(sizeof(double) * (size_t)N);
    double* x = (double*)malloc(sizeof(double) * (size_t)mn);
    if (!b || !ax || !x) { fprintf(stderr, "out of host memory\n"); rc = 2; goto done; }
    memcpy(b, pb.r, sizeof(double) * (size_t)n);                       /* b = [r ; l], Class1/APD_SsN_Class1.m:33 */
    memcpy(b + n, pb.l, sizeof(double) * (size_t)m);

    GPU(ssn_create(&ctx, 0));
    GPU(ssn_malloc(ctx, sizeof(double) * (size_t)mn, &c_d));
    GPU(ssn_malloc(ctx, sizeof(double) * (size_t)mn, &x_d));
    GPU(ssn_malloc(ctx, sizeof(double) * (size_t)N, &b_d));
    GPU(ssn_malloc(ctx, sizeof(double) * (size_t)N, &lk_d));
    GPU(ssn_malloc(ctx, sizeof(double) * (size_t)N, &ax_d));
    GPU(ssn_malloc(ctx, sizeof(double) * (size_t)m, &p_d));
    GPU(ssn_malloc(ctx, sizeof(double) * (size_t)n, &q_d));
    GPU(ssn_memcpy_h2d(ctx, c_d, pb.c, sizeof(double) * (size_t)mn));
    GPU(ssn_memcpy_h2d(ctx, b_d, b, sizeof(double) * (size_t)N));
    GPU(ssn_memcpy_h2d(ctx, p_d, pb.p, sizeof(double) * (size_t)m));
    GPU(ssn_memcpy_h2d(ctx, q_d, pb.q, sizeof(double) * (size_t)n));
    if (!all_inf(pb.gama, mn)) {
        GPU(ssn_malloc(ctx, sizeof(double) * (size_t)mn, &g_d));
        GPU(ssn_memcpy_h2d(ctx, g_d, pb.gama, sizeof(double) * (size_t)mn));
    }
    GPU(ssn_warmup_class1(ctx, (const double*)c_d, (const double*)b_d, (const double*)p_d, (const double*)q_d, m, n,
                          (const double*)g_d, INFINITY, maxit, (double*)x_d, (double*)lk_d));
    GPU(ssn_ax(ctx, (const double*)x_d, (const double*)p_d, (const double*)q_d, m, n, (double*)ax_d));
    GPU(ssn_memcpy_d2h(ctx, ax, ax_d, sizeof(double) * (size_t)N));
    GPU(ssn_memcpy_d2h(ctx, x, x_d, sizeof(double) * (size_t)mn));
    {
        double num = 0, den = 0, cx = 0;
        for (int64_t i = 0; i < N; ++i) { num += (ax[i] - b[i]) * (ax[i] - b[i]); den += b[i] * b[i]; }
        for (int64_t i = 0; i < mn; ++i) cx += pb.c[i] * x[i];
        printf("warm start, %d iterations: ||Ax(xk) - b|| / ||b|| = %.3e, c'xk = %.10g, kernels launched: %lld\n",
               maxit, sqrt(num / den), cx, (long long)ssn_launch_count(ctx));
    }
done:
    if (ctx) {
        ssn_free(ctx, c_d); ssn_free(ctx, x_d); ssn_free(ctx, b_d); ssn_free(ctx, lk_d); ssn_free(ctx, ax_d);
        ssn_free(ctx, p_d); ssn_free(ctx, q_d); ssn_free(ctx, g_d);
        ssn_destroy(ctx);
    }
    free(b); free(ax); free(x);
    ssn_problem_free(&pb);
    return rc;
}
