/*
 * oracle/csrc/oracle_kernels.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Order-sensitive CPU kernels of the oracle, restated from the behaviour of the
 * MATLAB built-ins the reference calls (the MATLAB runtime itself is closed source
 * and absent; see DESIGN.md "parity unpinned").  Nothing under oracle/ is ever
 * called by the product path; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Frozen conventions:
 *   - sparse * sparse  = column-wise Gustavson, C(:,j) = sum_k A(:,k)*B(k,j) with k
 *     ascending over the stored entries of B(:,j); multiply then add, no FMA
 *     (compile with -ffp-contract=off).  Used where the reference writes
 *     `Pro'*A*Pro` (AMG/transfer.m:66) and `Dff\Affs*W1` (AMG/transfer.m:51).
 *   - rand             = mt19937ar, init_genrand(5489), genrand_res53
 *     (MATLAB start-up stream; AMG/mis_set.m:31,35, Hybrid_AMG.m:40,69).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- SpGEMM */

/* Pattern size of C = A*B (CSC, A is m x k, B is k x n).  Returns nnz(C)
 * counting structural entries (numerical zeros are dropped later by the caller). */
int64_t orc_spgemm_csc_symbolic(int64_t m, int64_t n,
                                const int64_t *Ap, const int64_t *Ai,
                                const int64_t *Bp, const int64_t *Bi,
                                int64_t *Cp)
{
    int64_t *mark = (int64_t *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int64_t));
    if (!mark) return -1;
    for (int64_t i = 0; i < m; ++i) mark[i] = -1;
    int64_t nnz = 0;
    Cp[0] = 0;
    for (int64_t j = 0; j < n; ++j) {
        for (int64_t pb = Bp[j]; pb < Bp[j + 1]; ++pb) {
            int64_t k = Bi[pb];
            for (int64_t pa = Ap[k]; pa < Ap[k + 1]; ++pa) {
                int64_t i = Ai[pa];
                if (mark[i] != j) { mark[i] = j; ++nnz; }
            }
        }
        Cp[j + 1] = nnz;
    }
    free(mark);
    return nnz;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* Numeric phase.  Row indices of every output column come out sorted ascending;
 * the value of C(i,j) is accumulated in ascending k, starting from the first
 * product (not from 0.0 + product; the two are bit-identical except for -0.0). */
int orc_spgemm_csc_numeric(int64_t m, int64_t n,
                           const int64_t *Ap, const int64_t *Ai, const double *Ax,
                           const int64_t *Bp, const int64_t *Bi, const double *Bx,
                           const int64_t *Cp, int64_t *Ci, double *Cx)
{
    double *acc = (double *)calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    int64_t *mark = (int64_t *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int64_t));
    if (!acc || !mark) { free(acc); free(mark); return -1; }
    for (int64_t i = 0; i < m; ++i) mark[i] = -1;
    for (int64_t j = 0; j < n; ++j) {
        int64_t top = Cp[j];
        for (int64_t pb = Bp[j]; pb < Bp[j + 1]; ++pb) {
            int64_t k = Bi[pb];
            double b = Bx[pb];
            for (int64_t pa = Ap[k]; pa < Ap[k + 1]; ++pa) {
                int64_t i = Ai[pa];
                double prod = Ax[pa] * b;
                if (mark[i] != j) { mark[i] = j; Ci[top++] = i; acc[i] = prod; }
                else              { acc[i] = acc[i] + prod; }
            }
        }
        qsort(Ci + Cp[j], (size_t)(top - Cp[j]), sizeof(int64_t), cmp_i64);
        for (int64_t pc = Cp[j]; pc < top; ++pc) Cx[pc] = acc[Ci[pc]];
    }
    free(acc); free(mark);
    return 0;
}

/* y = A*x for CSC A (column sweep, j ascending: y(i) += A(i,j)*x(j)); this is the
 * summation order of a sparse matrix times a dense vector in a CSC runtime. */
void orc_spmv_csc(int64_t m, int64_t n, const int64_t *Ap, const int64_t *Ai,
                  const double *Ax, const double *x, double *y)
{
    for (int64_t i = 0; i < m; ++i) y[i] = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        double xj = x[j];
        for (int64_t pa = Ap[j]; pa < Ap[j + 1]; ++pa) y[Ai[pa]] = y[Ai[pa]] + Ax[pa] * xj;
    }
}

/* ---------------------------------------------------------------- MT19937 */

typedef struct { uint32_t mt[624]; int mti; } orc_mt_t;

void orc_mt_init(orc_mt_t *s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->mti = 624;
}

static uint32_t orc_mt_next(orc_mt_t *s)
{
    if (s->mti >= 624) {
        uint32_t *mt = s->mt;
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->mti = 0;
    }
    uint32_t y = s->mt[s->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* genrand_res53: 53-bit resolution doubles in [0,1). */
void orc_mt_rand(orc_mt_t *s, int64_t count, double *out)
{
    for (int64_t i = 0; i < count; ++i) {
        uint32_t a = orc_mt_next(s) >> 5, b = orc_mt_next(s) >> 6;
        out[i] = ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
    }
}

int orc_mt_sizeof(void) { return (int)sizeof(orc_mt_t); }

/* ---------------------------------------------------------------- plan ops
 * Straight loops used as the single-thread CPU port when numpy's BLAS is not
 * wanted (cpu_baseline 'port' leg can use either; tests compare both). */

/* Ax.m:10-13 : y = [X'p ; Xq], X = reshape(x,m,n) column-major. */
void orc_ax(const double *x, const double *p, const double *q, int64_t m, int64_t n, double *y)
{
    for (int64_t i = 0; i < m; ++i) y[n + i] = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        const double *col = x + j * m;
        double cs = 0.0, qj = q[j];
        for (int64_t i = 0; i < m; ++i) {
            cs = cs + col[i] * p[i];
            y[n + i] = y[n + i] + col[i] * qj;
        }
        y[j] = cs;
    }
}

/* Aty.m:10-13 : z = vec(p*y1' + y2*q'). */
void orc_aty(const double *y, const double *p, const double *q, int64_t m, int64_t n, double *z)
{
    const double *y1 = y, *y2 = y + n;
    for (int64_t j = 0; j < n; ++j) {
        double *col = z + j * m;
        double y1j = y1[j], qj = q[j];
        for (int64_t i = 0; i < m; ++i) col[i] = p[i] * y1j + y2[i] * qj;
    }
}
