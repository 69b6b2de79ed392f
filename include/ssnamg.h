/*
 * ssnamg.h -- C ABI of libssnamg.so: the B200-native (sm_100a CUDA) semismooth-Newton
 * inner linear solve of the IPD-SsN-AMG optimal-transport solver.
 *
 * The reference has no FFI: its boundary is the set of MATLAB function names on the path
 * (resolved through `addpath`, Class1/APD_SsN_Class1.m:13-14).  Each entry point below is what
 * a MEX shim of the same name binds (see INTEGRATION.md and mex/), and cites the reference
 * function it replaces.  All reference paths are relative to the reference repository root.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ types, no exceptions cross this boundary.
 *   - every function returns an int status: SSN_OK (0) or a negative SSN_E_* code that maps
 *     1:1 onto a reference `error()` condition or a CUDA failure; ssn_last_error() has text.
 *   - pointers named *_dev are DEVICE pointers on the context's GPU; *_host are host pointers.
 *     Functions with the suffix _host take host buffers and do the host<->device copies
 *     themselves (what a MEX shim holding ordinary mxArrays calls).
 *   - the plan x = vec(X) is the column-major m x n transport plan (m contiguous), dual /
 *     right-hand-side vectors are [column part (n) ; row part (m)], p has length m, q length n.
 *   - sparse matrices are CSR with sorted column indices, 0-based int32 indices.  Every sparse
 *     matrix on this path is structurally symmetric, so the same arrays are the CSC form a
 *     MATLAB sparse matrix holds (after widening indices to mwIndex).
 *   - all floating point is IEEE fp64.  Setup arithmetic follows the frozen summation order of
 *     DESIGN.md so that sparsity patterns, strength graphs and C/F splittings are bit-exact.
 *   - entry points are synchronous at return (results visible to the caller) unless stated.
 */
#ifndef SSNAMG_H
#define SSNAMG_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SSN_API __attribute__((visibility("default")))
#else
#define SSN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status codes */
enum {
    SSN_OK                = 0,
    SSN_E_CUDA            = -1,  /* a CUDA runtime call failed (text in ssn_last_error)           */
    SSN_E_INVALID         = -2,  /* bad argument (null pointer, negative size, unknown option)     */
    SSN_E_PQ_ZERO         = -3,  /* "p or q contains 0 !!!!!"               Hybrid_AMG.m:18-19     */
    SSN_E_BIGPH_FNODE     = -4,  /* "amg_options.bigph = 1 requires Nf > 0" AMG/Class_AMG.m:36-40  */
    SSN_E_PCG_NF          = -5,  /* "SSOR for bigraph requires pcg_options.nf" PCG.m:64            */
    SSN_E_NOT_SQUARE      = -6,  /* "Adjacency matrix must be square"       components.m:33        */
    SSN_E_CF_PARTITION    = -7,  /* C/F split does not partition the nodes: AMG/transfer.m:46-47
                                    would index out of range in MATLAB                             */
    SSN_E_COARSEN_STALL   = -8,  /* no F node found: AMG/Class_AMG.m:76 would loop forever         */
    SSN_E_NOT_BIGRAPH     = -9,  /* A(1:Nf,1:Nf) is not diagonal (check commented out at
                                    AMG/Class_AMG.m:52-54)                                         */
    SSN_E_UNSUPPORTED     = -10, /* PCG precd 3/4, Hybrid_twogrid, ...              */
    SSN_E_TOO_LARGE       = -11, /* a sparse object exceeds the int32 index range                  */
    SSN_E_NOT_SPD         = -12, /* small-component Cholesky hit a non-positive pivot              */
    SSN_E_NO_HIERARCHY    = -13, /* MG_Vcycle/MG_Wcycle called without a live hierarchy (the
                                    reference reads cleared globals, AMG/Class_AMG.m:110)          */
    SSN_E_ASATZ_DIM       = -14  /* ASAtz.m:21 multiplies the m-by-n Q by p: needs m == n          */
};

typedef struct ssn_ctx ssn_ctx;

/* Device CSR matrix; arrays are owned by the context that produced them. */
typedef struct ssn_csr {
    int64_t  nrows, ncols, nnz;
    int32_t *rowptr_dev;   /* nrows+1 */
    int32_t *colidx_dev;   /* nnz, ascending inside every row */
    double  *val_dev;      /* nnz */
} ssn_csr;

/* amg_options (AMG/Class_AMG.m:20-34).  A negative / NaN field means "empty" -> the
 * reference's isempty() default is applied. */
typedef struct ssn_amg_options {
    double  retol;     /* default 1e-12 */
    int32_t bigph;     /* default 0     */
    int32_t maxit;     /* default 50    */
    double  theta;     /* default 1/4   */
    int32_t smoth;     /* default 3     */
    int32_t cycle;     /* 'v' (118) or 'w' (119); default 'v'; any other value runs no cycle,
                          like the reference's nargin==2 default `cycle = 1`                  */
    int32_t isnsp;     /* default 0     */
    int32_t inter;     /* default 1     */
    int32_t fnode;     /* required when bigph != 0 */
    const double *guess_dev;  /* NULL -> zeros */
} ssn_amg_options;

/* pcg_options (PCG.m:18-27). */
typedef struct ssn_pcg_options {
    double  retol;     /* default 1e-11 */
    int32_t maxit;     /* default 10000 */
    int32_t precd;     /* 1 none, 2 Jacobi (default), 3 SSOR, 4 ichol (IC(0)), 5 bi-SSOR (needs nf) */
    int32_t nf;        /* <= 0: absent */
    const double *guess_dev;  /* NULL -> zeros */
} ssn_pcg_options;

/* prob_data (Class1/APD_SsN_Class1.m:154-156, Class2/APD_SsN_Class2.m:163-166). */
typedef struct ssn_prob_data {
    double  bk1, tk;
    int64_t m, n;
    const double  *p_dev, *q_dev;   /* m, n */
    const double  *t_dev;           /* diag(T), n+m; NULL -> zeros */
    const ssn_csr *H0;              /* (n+m) x (n+m), from ssn_asat */
    const double  *z_dev;           /* rhs: n+m (n+m+1 for the POT entry points) */
    const uint8_t *s_dev;           /* POT only: logical active set, m*n */
    const double  *phi_dev;         /* POT only: m*n */
} ssn_prob_data;

/* ------------------------------------------------------------------ context */
SSN_API int  ssn_create(ssn_ctx **ctx, int device);           /* device < 0: current device */
SSN_API int  ssn_destroy(ssn_ctx *ctx);
/* The process-wide default context, reference counted: every caller of ssn_default_ctx_acquire gets the SAME
 * context (created on the current device by the first one), ssn_default_ctx_release drops one reference and
 * the last one destroys it.  This is what the MEX shims use (mex/ssn_mex_common.h): the reference keeps the AMG
 * hierarchy in MATLAB globals (`global Ack Prok J smoth_it Rk`, AMG/Class_AMG.m:43,110; AMG/MG_Vcycle.m:8;
 * AMG/MG_Wcycle.m:9; AMG/transfer.m:17) and draws from ONE global `rand` stream (AMG/mis_set.m:31,35;
 * Hybrid_AMG.m:40,69; Class1/APD_SsN_Class1.m:246), both shared by all functions of the process -- so
 * Class_AMG.mex, MG_Wcycle.mex, transfer.mex, mis_set.mex and Hybrid_AMG.mex must share one hierarchy handle
 * and one MT19937 stream, which live in this context.  libssnamg.so is mapped once per process however many
 * MEX files link it.  Thread-safe; the entry points themselves are not re-entrant (MATLAB calls MEX files on
 * its interpreter thread). */
SSN_API int  ssn_default_ctx_acquire(ssn_ctx **ctx);
SSN_API int  ssn_default_ctx_release(void);
SSN_API int  ssn_default_ctx_refcount(void);
SSN_API const char *ssn_last_error(ssn_ctx *ctx);
SSN_API int  ssn_set_stream(ssn_ctx *ctx, void *cuda_stream); /* cudaStream_t; NULL = default stream */
SSN_API int  ssn_synchronize(ssn_ctx *ctx);
SSN_API int  ssn_version(void);
/* kernels launched by this context since creation (bench.py's gpu_launches claim) */
SSN_API int64_t ssn_launch_count(ssn_ctx *ctx);

/* Phase profiler (development aid): when enabled, named phases of the solve are timed with the
 * stream synchronised on both sides; ssn_profile_dump returns a text table and resets it. */
SSN_API int  ssn_profile_enable(ssn_ctx *ctx, int on);
SSN_API const char *ssn_profile_dump(ssn_ctx *ctx);
/* CUDA-event timer around the launches of the plan-wide kernels (the fused residual kernel and the
 * batched line-search kernel), on the launching stream: ssn_kernel_timer(ctx, 1) switches it on and
 * resets it, ssn_kernel_timer_read returns the accumulated milliseconds and the number of launches.
 * Each timed launch is synchronised, so leave it off outside measurements (bench.py's roofline). */
SSN_API int  ssn_kernel_timer(ssn_ctx *ctx, int on);
SSN_API int  ssn_kernel_timer_read(ssn_ctx *ctx, double *total_ms, int64_t *launches);
/* Solver tuning knobs (also read from the environment at ssn_create: SSN_DENSE_TAIL, SSN_DENSE_MAXN):
 * dense_tail != 0 collapses the tail of small AMG levels (N <= dense_max_n) into dense cycle
 * operators (one matvec per visit); 0 walks them step by step.  dense_max_n <= 0 keeps the value. */
SSN_API int  ssn_set_dense_tail(ssn_ctx *ctx, int dense_tail, int dense_max_n);
/* on != 0 (default; env SSN_PERSIST): Class_AMG's solve loop runs as one persistent cooperative kernel
 * (grid barriers between the dependent steps); 0 launches it kernel by kernel. */
SSN_API int  ssn_set_persistent(ssn_ctx *ctx, int on);
/* on != 0 (default): PCG's SSOR / IC(0) factors (precd 3 / 4), their dependency levels and row groups are built on
 * the device; 0 (env SSN_DEVICE_SETUP=0): on the host, once per call (kept as the cross-check of the device path). */
SSN_API int  ssn_set_device_setup(ssn_ctx *ctx, int on);
/* on != 0 (env SSN_FUSED_SETUP=1): the levels of the Class_AMG hierarchy with N <= 4096 rows are coarsened by ONE
 * kernel (strength, MIS rounds, interpolation, Galerkin products, smoother data; sizes never leave the device);
 * 0 (default): kernel by kernel like the large levels.  Same hierarchy bit for bit; the one-CTA kernel is
 * latency-bound on a single SM and loses to the piecewise path on a B200, so it is opt-in. */
SSN_API int  ssn_set_fused_setup(ssn_ctx *ctx, int on);
/* on != 0 (default; env SSN_CLUSTER_SOLVE): Class_AMG's solve loop of a late-phase hierarchy (<= 2^20 nonzeros on the
 * explicit levels, env SSN_CLUSTER_MAXNNZ) runs inside ONE thread-block cluster with the level vectors in the cluster's
 * distributed shared memory; 1: inside one cluster with the vectors in global memory (env SSN_DSM_SOLVE=0); 0: always as
 * the grid-wide kernel. */
SSN_API int  ssn_set_cluster_solve(ssn_ctx *ctx, int on);
/* Sparse products (ssn_spgemm, the Galerkin products of the AMG setup) whose intermediate upper bound exceeds `limit`
 * entries are formed slab of rows by slab of rows and concatenated (same rows bit for bit).  Default and maximum 2^30:
 * what 32-bit offsets into the intermediate arrays allow; smaller values bound the workspace (12 bytes per entry). */
SSN_API int  ssn_set_spgemm_slab_limit(ssn_ctx *ctx, int64_t limit);
/* cycles per grid-wide barrier of the persistent solve kernel: which = 0 cooperative-groups grid.sync(),
 * 1 = the library's own barrier (development aid) */
SSN_API int  ssn_debug_barrier_bench(ssn_ctx *ctx, int iters, int which, double *cycles_per_barrier);
/* cycle counters of the small-level cycle kernel: out64[0..63] (development aid) */
SSN_API int  ssn_debug_cycles(ssn_ctx *ctx, unsigned long long *out64, int reset);
/* per (operation, level) cycle counters of the persistent / cluster-resident solve kernels, filled only by a
 * library built with -DSSN_PERSIST_DEBUG: out256[op*16 + level] cycles, out256[128 + op*16 + level] calls
 * (op 0 residual, 1 block Gauss-Seidel update, 2 Jacobi step, 3 restriction / prolongation, 4 dense tail, 7 kernel) */
SSN_API int  ssn_debug_cycles_persist(ssn_ctx *ctx, unsigned long long *out256, int reset);

/* The library-owned MATLAB random stream (mt19937ar, init_genrand(5489), genrand_res53):
 * stands for MATLAB's global `rand` state consumed at AMG/mis_set.m:31,35 and
 * Hybrid_AMG.m:40,69.  Generated on the device. */
SSN_API int  ssn_rng_reset(ssn_ctx *ctx, uint32_t seed);
SSN_API int64_t ssn_rng_drawn(ssn_ctx *ctx);
SSN_API int  ssn_rand(ssn_ctx *ctx, int64_t count, double *out_dev);

/* device memory helpers for hosts without their own allocator (MEX shim, ctypes tests) */
SSN_API int  ssn_malloc(ssn_ctx *ctx, size_t bytes, void **ptr_dev);
SSN_API int  ssn_free(ssn_ctx *ctx, void *ptr_dev);
SSN_API int  ssn_memcpy_h2d(ssn_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
SSN_API int  ssn_memcpy_d2h(ssn_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
SSN_API int  ssn_memcpy_d2d(ssn_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);
SSN_API int  ssn_csr_free(ssn_ctx *ctx, ssn_csr *A);
SSN_API int  ssn_csr_upload(ssn_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                    const int32_t *rowptr_host, const int32_t *colidx_host,
                    const double *val_host, ssn_csr *out);
SSN_API int  ssn_csr_download(ssn_ctx *ctx, const ssn_csr *A, int32_t *rowptr_host,
                      int32_t *colidx_host, double *val_host);

/* ------------------------------------------------------------------ L1: plan operators */

/* y = Ax(x,p,q)  -- Ax.m:10-13.  y_dev: n+m.  One pass over the plan. */
SSN_API int ssn_ax(ssn_ctx *ctx, const double *x_dev, const double *p_dev, const double *q_dev,
           int64_t m, int64_t n, double *y_dev);
SSN_API int ssn_ax_host(ssn_ctx *ctx, const double *x_host, const double *p_host, const double *q_host,
                int64_t m, int64_t n, double *y_host);

/* z = Aty(y,p,q) -- Aty.m:10-13.  z_dev: m*n. */
SSN_API int ssn_aty(ssn_ctx *ctx, const double *y_dev, const double *p_dev, const double *q_dev,
            int64_t m, int64_t n, double *z_dev);
SSN_API int ssn_aty_host(ssn_ctx *ctx, const double *y_host, const double *p_host, const double *q_host,
                 int64_t m, int64_t n, double *z_host);

/* Fused SsN residual pieces, one read of w (Class1/APD_SsN_Class1.m:139-140,144,184,193):
 *   z   = (1/tk) * (w - Aty(lam))            (rounded exactly like the reference expression)
 *   s   = (z >= 0) & (z <= gama)             -> s_out_dev   (uint8, m*n)       [optional]
 *   px  = min(max(0,z),gama)                 -> prox_out_dev (m*n)             [optional]
 *   Ax(px)                                   -> axp_out_dev  (n+m)             [optional]
 *   the line-search term of :183-187         -> *norm2_out (host)              [optional]
 *       gama = Inf (prob < 3): ||px||^2;  finite gama (prob = 3): ||z||^2 - ||z - px||^2 = sum px*(2z - px)
 *   nnz(s)                                   -> *count_out (host)              [optional]
 * gama_dev == NULL means gama = gama_scalar for every entry (Inf for plain OT). */
SSN_API int ssn_prox_residual(ssn_ctx *ctx, const double *w_dev, const double *lam_dev,
                      const double *p_dev, const double *q_dev, int64_t m, int64_t n,
                      double tk, const double *gama_dev, double gama_scalar,
                      double *axp_out_dev, double *prox_out_dev, double *z_out_dev,
                      uint8_t *s_out_dev, double *norm2_out, int64_t *count_out);

/* ssn_prox_residual without its host read: the norm term and nnz(s) are left in scal2_dev[0..1] (device), nothing is
 * synchronised -- for a caller that sends them on (the row-sharded step puts them into the message of its all-reduce). */
SSN_API int ssn_prox_residual_dev(ssn_ctx *ctx, const double *w_dev, const double *lam_dev, const double *p_dev,
                      const double *q_dev, int64_t m, int64_t n, double tk, const double *gama_dev, double gama_scalar,
                      double *axp_out_dev, double *prox_out_dev, double *z_out_dev, uint8_t *s_out_dev, double *scal2_dev);

/* The same for partial OT (Class2/APD_SsN_Class2.m:124-130, 137-150, 196-217; u = [x (m*n); y (n); z (m)], lk of n+m+1
 * entries, H = [A I; phi' 0]), one read of w and one of phi:
 *   zk  = (1/tk) * (wk - [Aty(lk(1:n+m)) + lk(n+m+1)*phi ; lk(1:n+m)])     (rounded like the reference expression)
 *   s   = zk(1:m*n) >= 0                         -> s_out_dev   (uint8, m*n)       [optional]
 *   t   = zk(m*n+1:end) >= 0                     -> t_out_dev   (n+m doubles, 0/1) [optional]
 *   prox(zk) = max(zk,0)                         -> prox_out_dev (m*n+n+m)         [optional]
 *   [Ax(prox x) + [prox y; prox z] ; phi'*prox x] -> hp_out_dev  (n+m+1)           [optional]
 *   ||prox(zk)||^2 -> *norm2_out, nnz(s) -> *count_out (host)                      [optional] */
SSN_API int ssn_prox_residual_pot(ssn_ctx *ctx, const double *w_dev, const double *lam_dev, const double *p_dev,
                      const double *q_dev, int64_t m, int64_t n, double tk, const double *phi_dev,
                      double *hp_out_dev, double *prox_out_dev, uint8_t *s_out_dev, double *t_out_dev,
                      double *norm2_out, int64_t *count_out);

/* Batched line-search trials: n2_out_dev[t] = ||prox((w - Aty(lamT[t]))/tk)||^2 (finite gama: the prob = 3
 * term of ssn_prox_residual's norm2_out) for the nt <= 8
 * trial dual vectors lamT_dev[t*(n+m) .. ] in ONE read of w (Class1/APD_SsN_Class1.m:193-207
 * evaluates one trial per Aty + prox + norm pass).  Per-entry arithmetic as ssn_prox_residual. */
SSN_API int ssn_prox_trials(ssn_ctx *ctx, const double *w_dev, const double *lamT_dev, int nt,
                    const double *p_dev, const double *q_dev, int64_t m, int64_t n, double tk,
                    const double *gama_dev, double gama_scalar, double *n2_out_dev);

/* Screened batched trials of ONE search direction (gama = Inf): out_dev[t] = ||prox((w - Aty(lam +
 * delta^(ll0+t)*zeta))/tk)||^2 for t < nt <= 256 -- the values of ssn_prox_trials on the same trial
 * vectors up to the summation order -- and out_dev[nt] = number of entries at which some trial of
 * the batch may be active (out of m*n).  Entries whose first and last trial residuals are both
 * safely negative add exactly 0 and are dropped (the residual is linear in the step); the others are
 * listed and evaluated by a second, grid-balanced kernel, so where the trial plans are sparse the
 * pass is HBM-bound for any nt.  Deterministic.  lam_dev / zeta_dev: [column part (n) ; row part (m)]. */
SSN_API int ssn_prox_trials_lin(ssn_ctx *ctx, const double *w_dev, const double *lam_dev, const double *zeta_dev,
                        const double *p_dev, const double *q_dev, int64_t m, int64_t n, double tk,
                        double delta, int ll0, int nt, double *out_dev);

/* The O(m+n) half of a batch of Armijo trials (row-sharded callers use it with ssn_prox_trials on
 * their slab): lamT_out_dev[t] = lam + delta^(ll0+t)*zeta for t < nt <= 256 and
 * f0_out_dev[2t] = ||lamT[t]||^2, f0_out_dev[2t+1] = wlk'*lamT[t]  (Class1/APD_SsN_Class1.m:189-190). */
SSN_API int ssn_trial_vectors(ssn_ctx *ctx, const double *lam_dev, const double *zeta_dev, const double *wlk_dev,
                      int64_t N, double delta, int ll0, int nt, double *lamT_out_dev, double *f0_out_dev);

/* The Armijo backtracking of Class1/APD_SsN_Class1.m:182-211 (Class2 :170-200 with its own cF):
 *   lk_new = lk_old + delta^ll*zeta;  cF_new = bk1/2*||lk_new||^2 - wlk'*lk_new + tk/2*||prox(z)||^2;
 *   the first ll with !(cF_new > cF_old - nu*delta^ll*ress), or ll == ll_max, is accepted.
 * The first read of w evaluates ll = 0 alone (most steps accept it); every later read evaluates
 * `batch` (1..8) backtracking steps at once; batch <= 0 = adaptive: 8 to 128 steps per read through
 * the screened kernels of ssn_prox_trials_lin, by the measured sparsity of the trial plans.  Outputs: lam_new_dev (n+m), *ll_out,
 * *norm2_out = ||prox(z(lk_new))||^2, *cF_out, *passes_out = number of reads of w. */
SSN_API int ssn_linesearch(ssn_ctx *ctx, const double *w_dev, const double *lam_old_dev, const double *zeta_dev,
                   const double *wlk_dev, const double *p_dev, const double *q_dev, int64_t m, int64_t n,
                   double tk, double bk1, const double *gama_dev, double gama_scalar, double nu, double delta,
                   int ll_max, double cF_old, double ress, int batch, double *lam_new_dev, int *ll_out,
                   double *norm2_out, double *cF_out, int *passes_out);

/* The reference's Class 1 SCRIPT as one entry point: the warm start
 * (Class1/warmup_class1.m), the APD outer loop (Class1/APD_SsN_Class1.m:101-275) and the semismooth-Newton inner
 * loop (:137-238) with the inner linear solve selected by `inner_solver` (:66-70: 2 = PCG on Jk, 3 = aug_PCG,
 * 4 = Hybrid_AMG, 5 = Hybrid_twogrid), the Armijo line search (:182-211), the KKT bookkeeping (:239-274) and the
 * restart (:245-249) -- every plan-sized array stays on the device, nothing runs in an interpreter between the
 * kernels.  A zero / negative option field means "the script's value" (maxit 100, KKT_Tol 1e-6, 100 warm-start
 * iterations, inner_solver 4, amg_options / pcg_options of :81,:87-88). */
typedef struct ssn_apd_options {
    int32_t inner_solver;            /* 2, 3, 4 (default) or 5 */
    int32_t maxit;                   /* outer iterations, default 100 (:35) */
    double  KKT_Tol;                 /* default 1e-6 (:35) */
    int32_t warm_maxit;              /* A-ADMM iterations, default 100 (:59); < 0 = default, 0 = none */
    int32_t max_outer;               /* > 0: stop after this many outer iterations (tests / traces) */
    double  max_seconds;             /* > 0: stop once the outer loop ran this long */
    int32_t verbose;                 /* != 0: the script's progress lines on stderr */
    const ssn_amg_options *amg;      /* NULL: Class1/APD_SsN_Class1.m:87-88 */
    const ssn_pcg_options *pcg;      /* NULL: :81 */
} ssn_apd_options;
typedef struct ssn_apd_result {
    int32_t outer_its, converged;
    double  rel_kkt, objective;
    int32_t ssn_steps, ls_trials, ls_passes, amg_calls;
    double  warmup_s, loop_s, solve_s, asat_s, plan_s;     /* host wall clock: warm start, outer loop, and inside it
                                                              the inner solves / ASAt / the plan-wide part of the SsN steps */
    int32_t hist_len;                /* outer_its + 1 entries written to the history buffers */
    int64_t steps_len;               /* SsN steps taken (7 doubles each in steps_host, up to steps_cap) */
} ssn_apd_result;
/* c, r, l, p, q (and gama_dev, or NULL + gama_scalar) on the device; xk_out_dev (m*n), lk_out_dev (n+m) on the
 * device.  Optional HOST buffers: fxk / KKT_xk / KKT_lk histories (maxit+1 doubles each), SsN steps per outer
 * iteration (maxit ints), and per SsN step {k, ssn_it, nnz(s), components, inner iterations, ll, |F|}. */
SSN_API int ssn_apd_ssn_class1(ssn_ctx *ctx, const double *c_dev, const double *r_dev, const double *l_dev,
                       const double *p_dev, const double *q_dev, int64_t m, int64_t n, const double *gama_dev,
                       double gama_scalar, const ssn_apd_options *opts, double *xk_out_dev, double *lk_out_dev,
                       ssn_apd_result *result, double *fxk_hist_host, double *kkt_xk_hist_host,
                       double *kkt_lk_hist_host, int32_t *ssn_its_hist_host, double *steps_host, int64_t steps_cap);
/* The same solve for a caller that holds HOST arrays (a MATLAB script without gpuArray): inputs are copied to
 * the device once, the plan xk (m*n) and the duals lk (n+m) are copied back once -- one plugin call per solve
 * instead of one host<->device round trip of plan-sized arrays per operator call. */
SSN_API int ssn_apd_ssn_class1_host(ssn_ctx *ctx, const double *c_host, const double *r_host, const double *l_host,
                            const double *p_host, const double *q_host, int64_t m, int64_t n, const double *gama_host,
                            double gama_scalar, const ssn_apd_options *opts, double *xk_out_host, double *lk_out_host,
                            ssn_apd_result *result, double *fxk_hist_host, double *kkt_xk_hist_host,
                            double *kkt_lk_hist_host, int32_t *ssn_its_hist_host, double *steps_host, int64_t steps_cap);

/* ONE semismooth-Newton step of the script (Class1/APD_SsN_Class1.m:137-212) at a fixed APD state -- wk (m*n), wlk (n+m),
 * bk1, tk as :120-126 leave them -- from the duals lk: fused residual + active set -> ASAt -> Hybrid_AMG (inner_solver 4)
 * or Hybrid_twogrid (5) -> the Armijo loop -> the new residual.  Outputs lk_new, Fk_new (n+m each) and, optionally, 12
 * doubles on the HOST: nnz(s), nnz(H0), components, it_num, inner iterations, its relative residual, accepted ll, reads
 * of wk by the line search, ||F(lk)||, ||F(lk_new)||, ms of the plan-wide part, ms of the inner solve.
 * amg == NULL: the options of :87-88.  The _host variant takes and returns HOST arrays (one plugin call per step for a
 * MATLAB caller without gpuArray; wk travels over PCIe every call). */
SSN_API int ssn_ssn_step_class1(ssn_ctx *ctx, const double *wk_dev, const double *lk_dev, const double *wlk_dev,
                        const double *p_dev, const double *q_dev, int64_t m, int64_t n, double bk1, double tk,
                        const double *gama_dev, double gama_scalar, int inner_solver, const ssn_amg_options *amg,
                        double *lk_new_dev, double *Fk_new_dev, double *info12_host);
SSN_API int ssn_ssn_step_class1_host(ssn_ctx *ctx, const double *wk_host, const double *lk_host, const double *wlk_host,
                        const double *p_host, const double *q_host, int64_t m, int64_t n, double bk1, double tk,
                        const double *gama_host, double gama_scalar, int inner_solver, const ssn_amg_options *amg,
                        double *lk_new_host, double *Fk_new_host, double *info12_host);

/* ------------------------------------------------------------------ Class 2 (partial OT) as single calls
 * u = [x (m*n, column-major) ; y (n) ; z (m)], b = [r ; l ; mu], duals lk of n+m+1 entries, H = [A I ; phi' 0]. */

/* [uk,lk] = warmup_class2(c,r,l,p,q,mu,phi,0,maxit) -- Class2/warmup_class2.m:18-108, the A-ADMM warm start called at
 * Class2/APD_SsN_Class2.m:50, device resident: two fused plan-wide kernels per iteration (phi rides along as one more
 * streamed array), the slack blocks and the (n+m+1)-vectors in one block, invHHt (Class2/invHHt.m) in one block.
 * b_dev = [r ; l ; mu] (n+m+1).  Outputs uk_out_dev (m*n+n+m) and lk_out_dev (n+m+1).  The reference's stopping test is
 * commented out (warmup_class2.m:89-102), so exactly maxit iterations run. */
SSN_API int ssn_warmup_class2(ssn_ctx *ctx, const double *c_dev, const double *b_dev, const double *p_dev, const double *q_dev,
                      int64_t m, int64_t n, const double *phi_dev, int maxit, double *uk_out_dev, double *lk_out_dev);

/* Class2/APD_SsN_Class2.m:121-122 in one pass over the x block: wk = -wc + bk*(uk+ak*vk)/ak^2 (m*n+n+m),
 * huk = [Ax(xk)+[yk;zk] ; phi'*xk] and wlk = bk1*(lk - 1/bk*(huk - b)) - b (n+m+1 each). */
SSN_API int ssn_apd_begin_pot(ssn_ctx *ctx, const double *c_dev, const double *uk_dev, const double *vk_dev, const double *p_dev,
                      const double *q_dev, int64_t m, int64_t n, const double *phi_dev, const double *b_dev, const double *lk_dev,
                      double ak, double bk, double bk1, double *wk_out_dev, double *huk_out_dev, double *wlk_out_dev);
/* Class2/APD_SsN_Class2.m:231-238 in one pass: uk1 = prox(zk) at the duals lk, vk1 = uk1 + (uk1-uk)/ak, huk1 = H*uk1 and, on
 * the HOST, scal5 = { c'*xk1, ||xk1-max(xk1-c-(Aty(lk)+lk(end)*phi),0)||^2, ||yk1-max(yk1-lk(1:n),0)||^2,
 * ||zk1-max(zk1-lk(n+1:n+m),0)||^2, ||huk1-b||^2 } (the squares of KKT_xk, KKT_yk, KKT_zk, KKT_lk). */
SSN_API int ssn_apd_end_pot(ssn_ctx *ctx, const double *c_dev, const double *wk_dev, const double *uk_dev, const double *lk_dev,
                    const double *p_dev, const double *q_dev, int64_t m, int64_t n, const double *phi_dev, const double *b_dev,
                    double tk, double ak, double *uk1_dev, double *vk1_dev, double *huk1_out_dev, double *scal5_host);

/* The script Class2/APD_SsN_Class2.m:25-285 as ONE call: warm start, APD outer loop, SsN inner loop with its Armijo line
 * search and KKT bookkeeping, restart rule (:253-257).  opts as ssn_apd_ssn_class1 (inner_solver 3 = PCG4POT, 4 = AMG4POT
 * (default), 5 = AMG4POT with str = 'twogrid'; opts->amg NULL: the options of :80-81).  Outputs on the device: uk (m*n+n+m),
 * lk (n+m+1).  Optional HOST buffers: fxk history (maxit+1), the four KKT histories interleaved {x,y,z,l} (4*(maxit+1)),
 * SsN steps per outer iteration (maxit ints), per SsN step {k, ssn_it, nnz(s), components, inner iterations, ll, |F|}. */
SSN_API int ssn_apd_ssn_class2(ssn_ctx *ctx, const double *c_dev, const double *r_dev, const double *l_dev, const double *p_dev,
                       const double *q_dev, int64_t m, int64_t n, double mu, const double *phi_dev, const ssn_apd_options *opts,
                       double *uk_out_dev, double *lk_out_dev, ssn_apd_result *result, double *fxk_hist_host,
                       double *kkt4_hist_host, int32_t *ssn_its_hist_host, double *steps_host, int64_t steps_cap);
/* The same for a caller that holds HOST arrays (inputs copied to the device once, uk and lk copied back once). */
SSN_API int ssn_apd_ssn_class2_host(ssn_ctx *ctx, const double *c_host, const double *r_host, const double *l_host,
                            const double *p_host, const double *q_host, int64_t m, int64_t n, double mu, const double *phi_host,
                            const ssn_apd_options *opts, double *uk_out_host, double *lk_out_host, ssn_apd_result *result,
                            double *fxk_hist_host, double *kkt4_hist_host, int32_t *ssn_its_hist_host, double *steps_host,
                            int64_t steps_cap);

/* ONE semismooth-Newton step of Class2/APD_SsN_Class2.m:137-217 at a fixed APD state -- wk (m*n+n+m), wlk (n+m+1), bk1, tk as
 * :116-122 leave them -- from the duals lk (n+m+1): fused residual + active flags -> ASAt -> AMG4POT (inner_solver 4; 5: its
 * 'twogrid' variant; 3: PCG4POT) -> the Armijo loop (one fused pass over wk and phi per trial) -> the new residual.
 * Outputs lk_new, Fk_new (n+m+1 each) and the 12 HOST doubles of ssn_ssn_step_class1.  amg / pcg NULL: :80-81 / :74. */
SSN_API int ssn_ssn_step_class2(ssn_ctx *ctx, const double *wk_dev, const double *lk_dev, const double *wlk_dev, const double *p_dev,
                        const double *q_dev, int64_t m, int64_t n, double bk1, double tk, const double *phi_dev, int inner_solver,
                        const ssn_amg_options *amg, const ssn_pcg_options *pcg, double *lk_new_dev, double *Fk_new_dev,
                        double *info12_host);
SSN_API int ssn_ssn_step_class2_host(ssn_ctx *ctx, const double *wk_host, const double *lk_host, const double *wlk_host,
                        const double *p_host, const double *q_host, int64_t m, int64_t n, double bk1, double tk,
                        const double *phi_host, int inner_solver, const ssn_amg_options *amg, const ssn_pcg_options *pcg,
                        double *lk_new_host, double *Fk_new_host, double *info12_host);

/* [xk,lk] = warmup_class1(c,r,l,p,q,gama,0,maxit) -- Class1/warmup_class1.m:18-96, the A-ADMM warm start
 * called at Class1/APD_SsN_Class1.m:59, device resident: two fused plan-wide kernels per iteration
 * (17 plan-sized reads/writes instead of the ~45 of the Ax/Aty/prox/vector-update chain).
 * b_dev = [r ; l] (n+m).  Outputs xk_out_dev (m*n) and lk_out_dev (n+m).  The reference's stopping
 * test is commented out (warmup_class1.m:83-91), so exactly maxit iterations run. */
SSN_API int ssn_warmup_class1(ssn_ctx *ctx, const double *c_dev, const double *b_dev, const double *p_dev,
                      const double *q_dev, int64_t m, int64_t n, const double *gama_dev, double gama_scalar,
                      int maxit, double *xk_out_dev, double *lk_out_dev);

/* One fused stage of a warm-start iteration, for callers that own the loop (the row-sharded driver runs it on
 * its row slab and exchanges the column sums between the stages; single-GPU callers use ssn_warmup_class1).
 * Plan-sized arrays are updated in place; dual vectors are in the slab's own form [column part (n) ; row part
 * of the slab's m rows].  Scalars ak, bk, gk of the iteration (Class1/warmup_class1.m:59-62).
 *   stage 0 (:63-67)  reads xk vk wk pik lk2 c, lk1, axk = Ax(xk), b; writes dd; out1 = Ax(dd)
 *   stage 1 (:70-75)  reads dd xk wk pik lk2, y = invAAt(Ax(dd)); rewrites xk vk wk pik lk2;
 *                     out1 = Ax(vk1), out2 = Ax(xk1)
 * out1 / out2 (n+m): column sums over the slab's rows (partial when sharded), then the rows' own sums. */
SSN_API int ssn_warm_stage(ssn_ctx *ctx, int stage, double *xk_dev, double *vk_dev, double *wk_dev, double *pik_dev,
                   double *lk2_dev, double *dd_dev, const double *c_dev, const double *p_dev, const double *q_dev,
                   const double *b_dev, const double *lk1_dev, const double *axk_dev, const double *y_dev,
                   int64_t m, int64_t n, const double *gama_dev, double gama_scalar, double ak, double bk,
                   double gk, double *out1_dev, double *out2_dev);

/* The plan-wide lines of the APD outer iteration around the SsN solve, fused:
 *   ssn_apd_begin  wk = -c + bk*(xk+ak*vk)/ak^2 and axk = Ax(xk)            Class1/APD_SsN_Class1.m:125-126
 *   ssn_apd_end    xk1 = prox((wk-Aty(lam))/tk), vk1 = xk1+(xk1-xk)/ak, axk1 = Ax(xk1),
 *                  *cx_out = c'*xk1, *kx2_out = ||xk1-prox(xk1-c-Aty(lam))||^2   :239-254
 * one read of each input, one write of each output. */
SSN_API int ssn_apd_begin(ssn_ctx *ctx, const double *c_dev, const double *xk_dev, const double *vk_dev,
                  const double *p_dev, const double *q_dev, int64_t m, int64_t n, double ak, double bk,
                  double *wk_out_dev, double *axk_out_dev);
SSN_API int ssn_apd_end(ssn_ctx *ctx, const double *c_dev, const double *wk_dev, const double *xk_dev,
                const double *lam_dev, const double *p_dev, const double *q_dev, int64_t m, int64_t n,
                double tk, double ak, const double *gama_dev, double gama_scalar, double *xk1_out_dev,
                double *vk1_out_dev, double *axk1_out_dev, double *cx_out, double *kx2_out);

/* H = ASAt(s,p,q) -- ASAt.m:14-19.  s: logical m*n (1 byte per entry).  H is
 * (n+m) x (n+m), column nodes first, explicit zeros dropped. */
SSN_API int ssn_asat(ssn_ctx *ctx, const uint8_t *s_dev, const double *p_dev, const double *q_dev,
             int64_t m, int64_t n, ssn_csr *H_out);
SSN_API int ssn_asat_host(ssn_ctx *ctx, const uint8_t *s_host, const double *p_host,
                  const double *q_host, int64_t m, int64_t n, ssn_csr *H_out);

/* Row-sharded (multi-GPU) form of ASAt.m:15: ssn_active_coo compacts the logical active set of a
 * row slab (m_loc x n, rows [row_offset, row_offset+m_loc) of the m_global x n plan) into global
 * column-major linear indices (what MATLAB's find(s) returns, 0-based; *lin_out is a device
 * array of *E_out entries, release it with ssn_free); ssn_asat_coo assembles H from the
 * ascending union of those lists -- O(E) integers cross NVLink instead of m*n bytes. */
SSN_API int ssn_active_coo(ssn_ctx *ctx, const uint8_t *s_dev, int64_t m_loc, int64_t n, int64_t row_offset,
                   int64_t m_global, int64_t **lin_out, int64_t *E_out);
SSN_API int ssn_asat_coo(ssn_ctx *ctx, const int64_t *lin_sorted_dev, int64_t E, const double *p_dev,
                 const double *q_dev, int64_t m, int64_t n, ssn_csr *H_out);

/* y = ASAtz(z,s,p,q) -- ASAtz.m:15-22, reproduced as written (Q*p at :21; needs m == n). */
SSN_API int ssn_asatz(ssn_ctx *ctx, const double *z_dev, const uint8_t *s_dev, const double *p_dev,
              const double *q_dev, int64_t m, int64_t n, double *y_dev);

/* y = invAAt(x,p,q,sg1,sg2) -- invAAt.m:7-20 (callers resolve the nargin defaults). */
SSN_API int ssn_invaat(ssn_ctx *ctx, const double *x_dev, const double *p_dev, const double *q_dev,
               int64_t m, int64_t n, double sg1, double sg2, double *y_dev);
/* y = invHHt(v,p,q,sg,phi) -- Class2/invHHt.m:7-17.  v, y: n+m+1. */
SSN_API int ssn_invhht(ssn_ctx *ctx, const double *v_dev, const double *p_dev, const double *q_dev,
               int64_t m, int64_t n, double sg, const double *phi_dev, double *y_dev);

/* ------------------------------------------------------------------ L2: AMG setup */

/* S = strength(A,which) -- AMG/strength.m:6-18.  S has the pattern of the nonzero
 * off-diagonal entries of A whose strength value is nonzero. */
SSN_API int ssn_strength(ssn_ctx *ctx, const ssn_csr *A, int which, ssn_csr *S_out);

/* [isC,isF,As] = mis_set(A,theta) -- AMG/mis_set.m:8-67.  isC/isF: uint8[N] on the device;
 * As_out (optional) is the logical strength matrix as a CSR with unit values.  Consumes the
 * context's random stream exactly like the reference consumes `rand`. */
SSN_API int ssn_mis_set(ssn_ctx *ctx, const ssn_csr *A, double theta, uint8_t *isC_dev,
                uint8_t *isF_dev, ssn_csr *As_out);

/* [indC,indF] = cf_split(S) -- AMG/cf_split.m:6-15 (S logical; its third output, a MATLAB
 * graph object, is rebuilt by the .m wrapper from S). */
SSN_API int ssn_cf_split(ssn_ctx *ctx, const ssn_csr *S, uint8_t *indC_dev, uint8_t *indF_dev);

/* [Ac,Pro,As,indC] = transfer(A,amg_options) -- AMG/transfer.m:8-66.  `level_J` stands for
 * the reference's `global J` (transfer.m:17).  As_out / indC_dev are optional. */
SSN_API int ssn_transfer(ssn_ctx *ctx, const ssn_csr *A, const ssn_amg_options *opts, int level_J,
                 ssn_csr *Ac_out, ssn_csr *Pro_out, ssn_csr *As_out, uint8_t *indC_dev);

/* Setup phase of Class_AMG (AMG/Class_AMG.m:41-85): builds the hierarchy that the reference
 * keeps in the globals Ack/Prok/Rk/J/smoth_it into the context-owned handle.  *levels_out
 * receives J.  ssn_amg_level() exposes level k (1-based): A_k, Pro_k (NULL for k=1). */
SSN_API int ssn_amg_setup(ssn_ctx *ctx, const ssn_csr *A, const ssn_amg_options *opts, int *levels_out);
SSN_API int ssn_amg_level(ssn_ctx *ctx, int k, ssn_csr *A_out, ssn_csr *Pro_out);
SSN_API int ssn_amg_clear(ssn_ctx *ctx);

/* e = MG_Vcycle(r,isnsp,k) / e = MG_Wcycle(r,isnsp,k,e) -- AMG/MG_Vcycle.m, AMG/MG_Wcycle.m,
 * on the context's live hierarchy.  e_dev is in/out for the W-cycle (initial guess, pass
 * zeros for the reference's default); k is 1-based. */
SSN_API int ssn_mg_vcycle(ssn_ctx *ctx, const double *r_dev, int isnsp, int k, double *e_dev);
SSN_API int ssn_mg_wcycle(ssn_ctx *ctx, const double *r_dev, int isnsp, int k, double *e_dev);

/* [x,it,rel_res,rel_resk,rhok] = Class_AMG(A,b,amg_options) -- AMG/Class_AMG.m:20-110.
 * rel_resk_host / rhok_host: caller buffers of maxit+1 doubles (optional), *hist_len_out
 * receives the number of valid entries.  The hierarchy is cleared on return (Class_AMG.m:110)
 * unless keep_hierarchy != 0. */
SSN_API int ssn_class_amg(ssn_ctx *ctx, const ssn_csr *A, const double *b_dev,
                  const ssn_amg_options *opts, int keep_hierarchy, double *x_dev, int *it_out,
                  double *rel_res_out, double *rel_resk_host, double *rhok_host,
                  int *hist_len_out);

/* [x,it,rel_res,rel_resk,rhok] = twogrid_bigph(A,b,amg_options) -- AMG/twogrid_bigph.m:24-116, the two-level
 * method behind Hybrid_twogrid: block Gauss-Seidel smoother and interpolation of the first Class_AMG
 * level (bigph), coarse correction by PCG(Ac, ., maxit 100, Jacobi).  amg_options.fnode is required;
 * the caller applies twogrid_bigph.m's own defaults (:14-22: nargin == 2 -> retol 1e-12, maxit 20,
 * smoth 10, isnsp 1; empty fields -> retol 0, maxit 50, smoth 3, isnsp 0) before the call -- fields
 * left "empty" here fall back to Class_AMG's.  History buffers as ssn_class_amg. */
SSN_API int ssn_twogrid_bigph(ssn_ctx *ctx, const ssn_csr *A, const double *b_dev,
                      const ssn_amg_options *opts, double *x_dev, int *it_out, double *rel_res_out,
                      double *rel_resk_host, double *rhok_host, int *hist_len_out);

/* [x,it,rel_res,rel_resk,rhok] = twogrid(A,b,amg_options) -- AMG/twogrid.m:1-150, the two-grid method for a
 * general graph Laplacian: bigph = 1 is twogrid_bigph (fnode required, "bigph = 1 requires fnode > 0" :37
 * <-> SSN_E_BIGPH_FNODE, "Nf is not right for the bigraph" :47 <-> SSN_E_NOT_BIGRAPH); bigph = 0 smooths
 * with damped Jacobi 0.5*D^-1 and coarsens with mis_set(A,1/4) + standard interpolation (consumes the
 * random stream like transfer).  Defaults as twogrid_bigph (plus bigph 0), applied by the caller. */
SSN_API int ssn_twogrid(ssn_ctx *ctx, const ssn_csr *A, const double *b_dev, const ssn_amg_options *opts,
                double *x_dev, int *it_out, double *rel_res_out, double *rel_resk_host,
                double *rhok_host, int *hist_len_out);

/* ------------------------------------------------------------------ L2: Krylov */

/* [d,it,res,resk] = PCG(H,e,pcg_options) -- PCG.m:18-105.  resk_host: caller buffer of
 * maxit doubles (optional).  precd 3 (SSOR) and 4 (ichol, MATLAB's default IC(0); a nonpositive pivot is
 * SSN_E_NOT_SPD) apply their two sparse triangular solves level by level inside the persistent kernel;
 * the factors and their dependency levels are built once per call on the host. */
SSN_API int ssn_pcg(ssn_ctx *ctx, const ssn_csr *H, const double *e_dev, const ssn_pcg_options *opts,
            double *d_dev, int *it_out, double *res_out, double *resk_host);

/* ------------------------------------------------------------------ L3: dispatch */

/* [blocks,sizes,p,r] = components(A) -- components.m:32-55.  Component order: ascending
 * smallest member, members ascending (frozen convention).  blocks_dev: int32[N]
 * 1-based labels; p_dev: int32[N] 0-based; sizes/r are returned through the context:
 * *ncomp_out components; sizes_dev int32[N] (first ncomp valid), r_dev int32[N+1]. */
SSN_API int ssn_components(ssn_ctx *ctx, const ssn_csr *A, int32_t *blocks_dev, int32_t *sizes_dev,
                   int32_t *p_dev, int32_t *r_dev, int *ncomp_out);

/* [zeta,itamg,resamg,info] = Hybrid_AMG(prob_data,amg_options) -- Hybrid_AMG.m:11-113.
 * info_out[2] = {num_comp, it_num}. */
SSN_API int ssn_hybrid_amg(ssn_ctx *ctx, const ssn_prob_data *pd, const ssn_amg_options *opts,
                   double *zeta_dev, int *itamg_out, double *resamg_out, int *info_out);

/* [zeta,itamg,resamg,info] = Hybrid_twogrid(prob_data,amg_options) -- Hybrid_twogrid.m:11-89 (inner_solver = 5,
 * Class1/APD_SsN_Class1.m:178): Hybrid_AMG's dispatch with twogrid_bigph in place of Class_AMG. */
SSN_API int ssn_hybrid_twogrid(ssn_ctx *ctx, const ssn_prob_data *pd, const ssn_amg_options *opts,
                       double *zeta_dev, int *itamg_out, double *resamg_out, int *info_out);

/* [zeta,itpcg,respcg,info] = aug_PCG(prob_data,pcg_options) -- aug_PCG.m:11-37. */
SSN_API int ssn_aug_pcg(ssn_ctx *ctx, const ssn_prob_data *pd, const ssn_pcg_options *opts,
                double *zeta_dev, int *itpcg_out, double *respcg_out, int *info_out);

/* Class2/AMG4POT.m:27-55 and Class2/PCG4POT.m:26-40: bordered partial-OT solves.
 * pd->z_dev has n+m+1 entries, zeta_dev likewise. */
SSN_API int ssn_amg4pot(ssn_ctx *ctx, const ssn_prob_data *pd, const ssn_amg_options *opts,
                double *zeta_dev, int *it_out, double *res_out, int *info_out);
SSN_API int ssn_pcg4pot(ssn_ctx *ctx, const ssn_prob_data *pd, const ssn_pcg_options *opts,
                double *zeta_dev, int *it_out, double *res_out, int *info_out);
/* AMG4POT(prob_data,amg_options,str) with str = 'twogrid' when twogrid != 0 (Class2/AMG4POT.m:45-51: Hybrid_twogrid in
 * place of Hybrid_AMG; inner_solver = 5 of Class2/APD_SsN_Class2.m:181-182). */
SSN_API int ssn_amg4pot_str(ssn_ctx *ctx, const ssn_prob_data *pd, const ssn_amg_options *opts, int twogrid,
                double *zeta_dev, int *it_out, double *res_out, int *info_out);

/* The assembled rescaled system of Hybrid_AMG.m:17-24 / aug_PCG.m:16-22 (for parity tests):
 * Ae = bk1*Q0^2 + (K + Q0*H0*Q0)/tk and f = Q0*z. */
SSN_API int ssn_rescaled_system(ssn_ctx *ctx, const ssn_prob_data *pd, ssn_csr *Ae_out, double *f_dev);

/* Jk = bk1*speye(m+n) + (T + H0)/tk -- Class1/APD_SsN_Class1.m:147,151: the assembled KKT matrix that
 * inner_solver 1 / 2 hand to `\` / PCG (pd->t_dev may be NULL = T is zero; p, q, z are not read).  The diagonal of
 * Jk is always stored, also where H0 has none; `/tk` is a product with 1/tk (the oracle's frozen convention,
 * as in ssn_rescaled_system).  Release with ssn_csr_free. */
SSN_API int ssn_jk_system(ssn_ctx *ctx, const ssn_prob_data *pd, ssn_csr *Jk_out);

/* ------------------------------------------------------------------ sparse utilities
 * (generic building blocks of the path, exported for parity tests) */
SSN_API int ssn_spmv(ssn_ctx *ctx, const ssn_csr *A, const double *x_dev, double *y_dev);
SSN_API int ssn_spgemm(ssn_ctx *ctx, const ssn_csr *A, const ssn_csr *B, ssn_csr *C_out);
SSN_API int ssn_transpose(ssn_ctx *ctx, const ssn_csr *A, ssn_csr *At_out);

#ifdef __cplusplus
}
#endif
#endif /* SSNAMG_H */
