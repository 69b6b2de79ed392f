// amg_solve.cu -- solve phase: smoothers (block Gauss-Seidel on the bigraph level, damped
// Jacobi elsewhere, with the kernel-space correction), V/W-cycles, the coarse PCG, Class_AMG's
// stationary iteration and the general PCG of PCG.m.
//
// Large levels run as multi-block kernels (fused residual + partial sums, fused smoother
// update); the tail of small levels (N <= 2048) is ONE single-block kernel that walks the whole
// sub-hierarchy -- smoothing, restriction, the 2^(J-k) coarse PCG solves, prolongation -- with
// block barriers only, because that part of the W-cycle is pure launch latency otherwise.
#include "amg.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace ssn {

namespace {

constexpr int kCycleThreads = 1024;
constexpr int kMaxPartBlocks = 296;

// sum over a row's entries, cooperating TPR lanes; result valid in all TPR lanes.  The loop is
// batched 4 deep (all index/value loads, then all gathers, then the FMAs): the cores issue in
// order, so without this every iteration pays a full dependent load latency.
template <int TPR>
__device__ __forceinline__ double row_dot(const int* __restrict__ ptr, const int* __restrict__ idx,
                                          const double* __restrict__ val, const double* __restrict__ x,
                                          int row, int sub, bool valid) {
    double s = 0.0;
    if (valid) {
        const int e1 = ptr[row + 1];
        int e = ptr[row] + sub;
        for (; e + 3 * TPR < e1; e += 4 * TPR) {
            const int i0 = idx[e], i1 = idx[e + TPR], i2 = idx[e + 2 * TPR], i3 = idx[e + 3 * TPR];
            const double v0 = val[e], v1 = val[e + TPR], v2 = val[e + 2 * TPR], v3 = val[e + 3 * TPR];
            const double x0 = x[i0], x1 = x[i1], x2 = x[i2], x3 = x[i3];
            s = fma(v0, x0, s); s = fma(v1, x1, s); s = fma(v2, x2, s); s = fma(v3, x3, s);
        }
        for (; e < e1; e += TPR) s = fma(val[e], x[idx[e]], s);
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// increment of the smoother for row i:  coef + (R h)_i,  h = g - Axi*coef
template <int TPR>
__device__ __forceinline__ double smooth_inc(const LevelDev& L, const double* __restrict__ g, double coef, int post,
                                             int row, int sub, bool valid) {
    double hi = 0.0, di = 0.0;
    if (valid) { hi = g[row] - L.Axi[row] * coef; di = L.dinv[row]; }
    double s = 0.0;
    if (L.bigph) {
        // forward sweep (R, MG_Wcycle.m:19): row nodes see the updated column nodes;
        // backward sweep (R', MG_Wcycle.m:37): column nodes see the updated row nodes
        const bool coupled = valid && (post ? (row < L.Nf) : (row >= L.Nf));
        if (coupled) {
            const int e1 = L.ap[row + 1];
            for (int e = L.ap[row] + sub; e < e1; e += TPR) {
                const int j = L.ai[e];
                const bool other = post ? (j >= L.Nf) : (j < L.Nf);
                if (other) s = fma(L.av[e], L.dinv[j] * (g[j] - L.Axi[j] * coef), s);
            }
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    return coef + di * (hi - s);
}

// ------------------------------------------------------------------ multi-block kernels

// g = r - A*e (e == nullptr: g = r); part[2b] = sum g, part[2b+1] = sum g^2 over the block's rows
template <int TPR>
__global__ void __launch_bounds__(256) resid_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                                    const double* __restrict__ val, const double* __restrict__ r,
                                                    const double* __restrict__ e, double* __restrict__ g,
                                                    double* __restrict__ part) {
    __shared__ double red[32];
    const int sub = threadIdx.x % TPR;
    const int rows_per_pass = gridDim.x * (256 / TPR);
    double sg = 0.0, sg2 = 0.0;
    for (int base = 0; base < n; base += rows_per_pass) {
        const int row = base + blockIdx.x * (256 / TPR) + threadIdx.x / TPR;
        const bool valid = row < n;
        double d = 0.0;
        if (e != nullptr) d = row_dot<TPR>(ptr, idx, val, e, row, sub, valid);
        if (valid && sub == 0) {
            const double gi = r[row] - d;
            g[row] = gi; sg += gi; sg2 = fma(gi, gi, sg2);
        }
    }
    sg = block_sum(sg, red);
    sg2 = block_sum(sg2, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = sg; part[2 * blockIdx.x + 1] = sg2; }
}

__device__ __forceinline__ double sum_parts(const double* __restrict__ part, int np, int stride, double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) s += part[(size_t)i * stride];
    return block_sum(s, red);
}

// e += coef + R(g - Axi*coef), coef = isnsp ? sum(g)/xx : 0
template <int TPR>
__global__ void __launch_bounds__(256) smooth_apply_kernel(LevelDev L, const double* __restrict__ g, double* __restrict__ e,
                                                           const double* __restrict__ part, int np, int isnsp, int post,
                                                           int e_is_zero) {
    __shared__ double red[32];
    double coef = 0.0;
    if (isnsp) coef = sum_parts(part, np, 2, red) / L.xx;
    const int sub = threadIdx.x % TPR;
    const int rows_per_pass = gridDim.x * (256 / TPR);
    for (int base = 0; base < L.N; base += rows_per_pass) {
        const int row = base + blockIdx.x * (256 / TPR) + threadIdx.x / TPR;
        const bool valid = row < L.N;
        const double inc = smooth_inc<TPR>(L, g, coef, post, row, sub, valid);
        if (valid && sub == 0) e[row] = e_is_zero ? inc : (e[row] + inc);
    }
}

// One damped-Jacobi smoothing step with the kernel correction in ONE launch: the coefficient is
// (sum(r) - Axi'e)/xx, both sums arriving as block partials of earlier launches, and the update
// is row-local; ealt <- smoothed ecur, dot_out[b] <- partial of Axi'ealt for the next step.
template <int TPR>
__global__ void __launch_bounds__(256) smooth_fused_kernel(LevelDev L, const double* __restrict__ r,
                                                           const double* __restrict__ ecur, double* __restrict__ ealt,
                                                           const double* __restrict__ part_r, int np_r,
                                                           const double* __restrict__ dot_in, int np_d,
                                                           double* __restrict__ dot_out, int isnsp, int e_is_zero) {
    __shared__ double red[32];
    double coef = 0.0;
    if (isnsp) {
        const double sr = sum_parts(part_r, np_r, 1, red);
        const double sd = (np_d > 0) ? sum_parts(dot_in, np_d, 1, red) : 0.0;
        coef = (sr - sd) / L.xx;
    }
    const int sub = threadIdx.x % TPR;
    const int rows_per_pass = gridDim.x * (256 / TPR);
    double part = 0.0;
    for (int base = 0; base < L.N; base += rows_per_pass) {
        const int row = base + blockIdx.x * (256 / TPR) + threadIdx.x / TPR;
        const bool valid = row < L.N;
        double d = 0.0;
        if (!e_is_zero) d = row_dot<TPR>(L.ap, L.ai, L.av, ecur, row, sub, valid);
        if (valid && sub == 0) {
            const double axi = L.Axi[row];
            const double gi = r[row] - d;
            const double en = (e_is_zero ? 0.0 : ecur[row]) + coef + L.dinv[row] * (gi - axi * coef);
            ealt[row] = en;
            part = fma(axi, en, part);
        }
    }
    part = block_sum(part, red);
    if (threadIdx.x == 0) dot_out[blockIdx.x] = part;
}

// part[b] = sum over the block's slice of x[i]*(y ? y[i] : 1)
__global__ void __launch_bounds__(256) dot_parts_kernel(int n, const double* __restrict__ x, const double* __restrict__ y,
                                                        double* __restrict__ part) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += y ? x[i] * y[i] : x[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) reduce_parts_kernel(const double* __restrict__ part, int np, double* __restrict__ out) {
    __shared__ double red[32];
    const double a = sum_parts(part, np, 2, red);
    const double b = sum_parts(part + 1, np, 2, red);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

__global__ void axpy_kernel(int n, double a, const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}

// ------------------------------------------------------------------ single-block cycle
//
// One CTA of 1024 threads walks the whole sub-hierarchy.  Latency is everything here, so:
//   * a Jacobi smoothing step is ONE pass + ONE single-barrier reduction: with A symmetric,
//     sum(g) = sum(r) - (A*ones)'e, so the kernel-correction coefficient of step t is known
//     before its SpMV and the update e_i += coef + dinv_i*(g_i - Axi_i*coef) is row-local
//     (ping-pong between two buffers);
//   * the coarsest PCG (N <= 32) runs in a single warp with shuffles only;
//   * the second coarse solve at depth J-1 is skipped: it repeats the first one exactly
//     (same right-hand side, guess ignored, MG_Wcycle.m:30,44).

constexpr int kBT = 4;                                    // lanes per row inside the block kernel

// cycle counters of the single-block kernel (development aid; read through ssn_debug_cycles)
__device__ unsigned long long g_dbg_cycles[64];
#define DBG_T0() const long long dbg_t0 = clock64()
// per (operation, level) cycle counters of the persistent solve kernels: [op*16 + level] cycles, [128 + op*16 + level] calls
// (op: 0 resid, 1 gs_apply, 2 jacobi, 3 spmv, 4 dense, 7 whole kernel); built with -DSSN_PERSIST_DEBUG only
__device__ unsigned long long g_pdbg[256];
#ifdef SSN_PERSIST_DEBUG
#define PDBG(slot, call) do { const long long t0__ = clock64(); call; if (blockIdx.x == 0 && threadIdx.x == 0) { g_pdbg[((slot) - 25) * 16 + dbg_level__] += (unsigned long long)(clock64() - t0__); g_pdbg[128 + ((slot) - 25) * 16 + dbg_level__] += 1ull; } } while (0)
#else
#define PDBG(slot, call) do { call; } while (0)
#endif
#define DBG_ADD(slot) do { if (threadIdx.x == 0) { g_dbg_cycles[(slot)] += (unsigned long long)(clock64() - dbg_t0); g_dbg_cycles[32 + (slot)] += 1ull; } } while (0)

// block-wide sum with ONE barrier: warp partials go to the buffer selected by `flip`, which the
// caller alternates so that slow readers of call n never race with the writers of call n+1.
__device__ __forceinline__ double block_sum1(double v, double (*red)[32], int& flip) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    double* buf = red[flip & 1];
    if (lane == 0) buf[w] = v;
    __syncthreads();
    double t = buf[lane];                                 // kCycleThreads == 1024: exactly 32 warp partials
    t = warp_sum(t);
    ++flip;
    return t;
}

__device__ void blk_resid(const LevelDev& L, const double* r, const double* e, double* g) {
    const int sub = threadIdx.x % kBT;
    for (int base = 0; base < L.N; base += kCycleThreads / kBT) {
        const int row = base + threadIdx.x / kBT;
        const bool valid = row < L.N;
        double d = 0.0;
        if (e != nullptr) d = row_dot<kBT>(L.ap, L.ai, L.av, e, row, sub, valid);
        if (valid && sub == 0) g[row] = r[row] - d;
    }
    __syncthreads();
}

// Smoothing on level L; returns the buffer that holds e afterwards (ecur or ealt).
__device__ double* blk_smooth(const LevelDev& L, const double* r, double* ecur, double* ealt, bool e_zero, int steps,
                              int isnsp, int post, double sum_r, double (*red)[32], int& flip) {
    const int sub = threadIdx.x % kBT;
    if (steps == 0) {
        if (e_zero) { for (int i = threadIdx.x; i < L.N; i += kCycleThreads) ecur[i] = 0.0; __syncthreads(); }
        return ecur;
    }
    if (L.bigph) {                                        // block Gauss-Seidel level: residual pass + coupled update
        for (int it = 0; it < steps; ++it) {
            blk_resid(L, r, e_zero ? nullptr : ecur, L.g);
            double coef = 0.0;
            if (isnsp) {
                double sg = 0.0;
                for (int i = threadIdx.x; i < L.N; i += kCycleThreads) sg += L.g[i];
                coef = block_sum1(sg, red, flip) / L.xx;
            }
            for (int base = 0; base < L.N; base += kCycleThreads / kBT) {
                const int row = base + threadIdx.x / kBT;
                const bool valid = row < L.N;
                const double inc = smooth_inc<kBT>(L, L.g, coef, post, row, sub, valid);
                if (valid && sub == 0) ecur[row] = e_zero ? inc : (ecur[row] + inc);
            }
            e_zero = false;
            __syncthreads();
        }
        return ecur;
    }
    double dotAe = 0.0;
    if (isnsp && !e_zero) {
        double s = 0.0;
        for (int i = threadIdx.x; i < L.N; i += kCycleThreads) s = fma(L.Axi[i], ecur[i], s);
        dotAe = block_sum1(s, red, flip);
    }
    for (int it = 0; it < steps; ++it) {
        const double coef = isnsp ? (sum_r - dotAe) / L.xx : 0.0;
        double part = 0.0;
        for (int base = 0; base < L.N; base += kCycleThreads / kBT) {
            const int row = base + threadIdx.x / kBT;
            const bool valid = row < L.N;
            double d = 0.0;
            if (!e_zero) d = row_dot<kBT>(L.ap, L.ai, L.av, ecur, row, sub, valid);
            if (valid && sub == 0) {
                const double axi = L.Axi[row];
                const double gi = r[row] - d;
                const double en = (e_zero ? 0.0 : ecur[row]) + coef + L.dinv[row] * (gi - axi * coef);
                ealt[row] = en;
                part = fma(axi, en, part);
            }
        }
        dotAe = block_sum1(part, red, flip);              // the barrier inside also publishes ealt
        double* t = ecur; ecur = ealt; ealt = t;
        e_zero = false;
    }
    return ecur;
}

// y (+)= M*x for a CSR M with nrows rows
__device__ void blk_spmv(int nrows, const int* ptr, const int* idx, const double* val, const double* x, double* y, bool add) {
    const int sub = threadIdx.x % kBT;
    for (int base = 0; base < nrows; base += kCycleThreads / kBT) {
        const int row = base + threadIdx.x / kBT;
        const bool valid = row < nrows;
        const double d = row_dot<kBT>(ptr, idx, val, x, row, sub, valid);
        if (valid && sub == 0) y[row] = add ? (y[row] + d) : d;
    }
    __syncthreads();
}

__device__ __forceinline__ double level_diag(const LevelDev& L, int i) {
    return L.bigph ? (1.0 / L.dinv[i]) : (0.5 / L.dinv[i]);     // recover diag(A) from the smoother scaling
}

// PCG(A,r) with the defaults of PCG.m:18-23 (zero guess, retol 1e-11, maxit 1e4, Jacobi), N <= 32:
// one warp, lane i owns row i, vectors live in registers, SpMV gathers through shuffles.
__device__ void warp_pcg(const LevelDev& L, const double* rhs, double* x) {
    const int lane = threadIdx.x & 31;
    const bool valid = lane < L.N;
    const int e0 = valid ? L.ap[lane] : 0, e1 = valid ? L.ap[lane + 1] : 0;
    int maxlen = e1 - e0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    const double diag = valid ? level_diag(L, lane) : 1.0;
    double r = valid ? __ldcg(rhs + lane) : 0.0;
    double p = r / diag, xi = 0.0;
    double delta_new = warp_sum(r * p);
    const double delta_0 = delta_new, tol2 = 1e-11 * 1e-11;
    int it = 0;
    while (it < 10000 && delta_new > tol2 * delta_0) {
        const double delta_old = delta_new;
        double q = 0.0;
        for (int t = 0; t < maxlen; ++t) {
            const int e = e0 + t;
            const bool has = e < e1;
            const int j = has ? L.ai[e] : 0;
            const double a = has ? L.av[e] : 0.0;
            const double pj = __shfl_sync(0xffffffffu, p, j);
            q = fma(a, pj, q);
        }
        const double alpha = delta_old / warp_sum(q * p);
        xi += alpha * p;
        r -= alpha * q;
        const double w = r / diag;
        delta_new = warp_sum(r * w);
        p = w + (delta_new / delta_old) * p;
        ++it;
    }
    if (valid) x[lane] = xi;
}

// block-wide PCG for a coarsest level with more than 32 unknowns (other problem sizes)
__device__ void blk_pcg(const LevelDev& L, const double* rhs, double* x, double (*red)[32], int& flip) {
    const int n = L.N;
    double* r = L.pcg; double* p = r + n; double* q = p + n;
    double dn = 0.0;
    for (int i = threadIdx.x; i < n; i += kCycleThreads) {
        const double ri = rhs[i];
        const double pi = ri / level_diag(L, i);
        r[i] = ri; p[i] = pi; x[i] = 0.0; dn = fma(ri, pi, dn);
    }
    double delta_new = block_sum1(dn, red, flip);
    const double delta_0 = delta_new;
    const double tol2 = 1e-11 * 1e-11;
    int it = 0;
    while (it < 10000 && delta_new > tol2 * delta_0) {
        const double delta_old = delta_new;
        blk_spmv(n, L.ap, L.ai, L.av, p, q, false);
        double qp = 0.0;
        for (int i = threadIdx.x; i < n; i += kCycleThreads) qp = fma(q[i], p[i], qp);
        qp = block_sum1(qp, red, flip);
        const double alpha = delta_old / qp;
        double dnew = 0.0;
        for (int i = threadIdx.x; i < n; i += kCycleThreads) {
            x[i] += alpha * p[i];
            const double ri = r[i] - alpha * q[i];
            r[i] = ri; q[i] = ri / level_diag(L, i);      // q now holds w = M^{-1} r
            dnew = fma(ri, q[i], dnew);
        }
        delta_new = block_sum1(dnew, red, flip);
        const double beta = delta_new / delta_old;
        for (int i = threadIdx.x; i < n; i += kCycleThreads) p[i] = q[i] + beta * p[i];
        __syncthreads();
        ++it;
    }
    __syncthreads();
}

// The whole cycle from level k0 down, iteratively (phase machine instead of recursion).
__global__ void __launch_bounds__(kCycleThreads) coarse_cycle_kernel(const LevelDev* __restrict__ levels, int k0, int J,
                                                                     int smoth, int isnsp, int wcycle, int e0_zero) {
    __shared__ double red[2][32];
    __shared__ LevelDev sl[16];
    const int nl = J - k0;
    for (int t = threadIdx.x; t < nl && t < 16; t += kCycleThreads) sl[t] = levels[k0 + t];
    __syncthreads();
    int flip = 0;
    int phase[16]; bool zero[16]; double* ecur[16]; double* ealt[16]; double sum_r[16];
    for (int t = 0; t < 16; ++t) { ecur[t] = nullptr; ealt[t] = nullptr; sum_r[t] = 0.0; phase[t] = 0; zero[t] = true; }
    for (int t = 0; t < nl; ++t) { ecur[t] = sl[t].e; ealt[t] = sl[t].pcg; }
    int k = 0;
    zero[0] = e0_zero != 0;
    while (true) {
        const LevelDev& L = sl[k];
        if (k == nl - 1) {                                // coarsest: PCG(A,r), guess ignored (MG_Wcycle.m:44)
            { DBG_T0();
            if (L.N <= 32) { if (threadIdx.x < 32) warp_pcg(L, L.r, L.e); __syncthreads(); }
            else blk_pcg(L, L.r, L.e, red, flip);
            DBG_ADD(24 + 0); }
            ecur[k] = L.e;
            if (k == 0) break;
            --k; continue;
        }
        if (phase[k] == 0) {
            if (isnsp && !L.bigph) {
                double s = 0.0;
                for (int i = threadIdx.x; i < L.N; i += kCycleThreads) s += L.r[i];
                sum_r[k] = block_sum1(s, red, flip);
            }
            double* en;
            { DBG_T0();
            en = blk_smooth(L, L.r, ecur[k], ealt[k], zero[k], smoth, isnsp, 0, sum_r[k], red, flip);   // presmoothing
            DBG_ADD(k); }
            if (en != ecur[k]) { ealt[k] = ecur[k]; ecur[k] = en; }
            const LevelDev& Lc = sl[k + 1];
            { DBG_T0();
            blk_resid(L, L.r, ecur[k], L.g);
            blk_spmv(Lc.N, Lc.tp, Lc.ti, Lc.tv, L.g, Lc.r, false);                  // restriction
            DBG_ADD(8 + k); }
            phase[k] = 1; phase[k + 1] = 0; zero[k + 1] = true; ecur[k + 1] = Lc.e; ealt[k + 1] = Lc.pcg; ++k; continue;
        }
        if (phase[k] == 1 && wcycle && (k + 1 != nl - 1)) {                         // correction again
            phase[k] = 2; phase[k + 1] = 0; zero[k + 1] = false; ++k; continue;
        }
        {
            const LevelDev& Lc = sl[k + 1];
            { DBG_T0();
            blk_spmv(L.N, Lc.pp, Lc.pi, Lc.pv, ecur[k + 1], ecur[k], true);         // prolongation
            DBG_ADD(16 + k); }
            double* en;
            { DBG_T0();
            en = blk_smooth(L, L.r, ecur[k], ealt[k], false, smoth, isnsp, 1, sum_r[k], red, flip);   // postsmoothing
            DBG_ADD(k); }
            if (en != ecur[k]) { ealt[k] = ecur[k]; ecur[k] = en; }
        }
        if (k == 0) break;
        --k;
    }
    // the caller reads level k0's correction from its e buffer
    if (ecur[0] != sl[0].e) {
        for (int i = threadIdx.x; i < sl[0].N; i += kCycleThreads) sl[0].e[i] = ecur[0][i];
    }
}

// ------------------------------------------------------------------ dense tail
//
// The tail of small levels (N <= kDenseMaxN) is visited 2^k times per W-cycle and every visit is a
// chain of latency-bound sparse steps.  With a zero guess a cycle on level k is a LINEAR map of
// its right-hand side, so the tail is collapsed once per hierarchy into explicit dense operators
// B_k (N_k x N_k, row-major): the coarsest B is PCG(A, e_i) column by column (PCG.m defaults,
// exactly the solves MG_Wcycle.m:44 performs), and B_k for k < J-1 is ONE level of the cycle
// (MG_Wcycle.m:15-42: smoth pre-smoothing steps with the kernel correction, restriction, one or
// two coarse corrections through B_{k+1}, prolongation, smoth post-smoothing steps) applied to the
// unit vectors, C columns per CTA so the matrix is read once per C columns.  A visit is then one
// dense matvec, e = B_k r (zero guess) or e += B_k (r - A_k e) (second W visit, MG_Wcycle.m:30).
// The results differ from the step-by-step kernel by rounding only.

constexpr int kDT = 1024;                                 // threads of the build kernel
constexpr int kDTW = kDT / 32;

// row stride (in doubles) of the [row][column] multi-vector layout in shared memory: C + 1 keeps the
// gathers x[idx[e]][c] of a warp spread over the banks (a stride of C = 8 doubles maps them onto 2)
template <int C> struct MV { static constexpr int S = (C == 1) ? 1 : ((C % 2 == 0) ? C + 1 : C + 2); };   // an odd stride

template <int C>
__device__ __forceinline__ void block_sumC(double (&v)[C], double* red, int& flip) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* buf = red + (flip & 1) * (C * 32);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double t = warp_sum(v[c]);
        if (lane == 0) buf[c * 32 + w] = t;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = warp_sum(buf[c * 32 + lane]);    // kDT == 1024: 32 warp partials
    ++flip;
}

// acc[c] = sum_e val[e] * x[idx[e]][c] over a row, kBT lanes per row, x in shared memory as [n][C]
template <int C>
__device__ __forceinline__ void row_dot_multi(const int* __restrict__ ptr, const int* __restrict__ idx,
                                              const double* __restrict__ val, const double* x, int row, int sub,
                                              bool valid, double (&acc)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.0;
    if (valid) {
        const int e1 = ptr[row + 1];
        int e = ptr[row] + sub;
        for (; e + kBT < e1; e += 2 * kBT) {
            const int j0 = idx[e], j1 = idx[e + kBT];
            const double a0 = val[e], a1 = val[e + kBT];
            const double* x0 = x + (size_t)j0 * MV<C>::S; const double* x1 = x + (size_t)j1 * MV<C>::S;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fma(a1, x1[c], fma(a0, x0[c], acc[c]));
        }
        if (e < e1) {
            const int j0 = idx[e]; const double a0 = val[e]; const double* x0 = x + (size_t)j0 * MV<C>::S;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fma(a0, x0[c], acc[c]);
        }
    }
#pragma unroll
    for (int o = kBT / 2; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
}

// y[nrows][C] (+)= M * x[.][C], CSR M, both vectors in shared memory
template <int C>
__device__ void spmm_smem(int nrows, const int* ptr, const int* idx, const double* val, const double* x, double* y, int mode) {
    const int sub = threadIdx.x % kBT;                    // mode 0: y = Mx, 1: y += Mx, 2: y = y - Mx
    for (int base = 0; base < nrows; base += kDT / kBT) {
        const int row = base + threadIdx.x / kBT;
        const bool valid = row < nrows;
        double acc[C];
        row_dot_multi<C>(ptr, idx, val, x, row, sub, valid, acc);
        if (valid && sub == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                double* yy = y + (size_t)row * MV<C>::S + c;
                *yy = (mode == 0) ? acc[c] : (mode == 1 ? *yy + acc[c] : *yy - acc[c]);
            }
        }
    }
    __syncthreads();
}

// y[n][C] (+)= B * x[n][C], dense row-major B (global), warp per row
template <int C>
__device__ void dense_smem(int n, const double* __restrict__ B, const double* x, double* y, bool add) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int row = w; row < n; row += kDTW) {
        double acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.0;
        const double* Br = B + (size_t)row * n;
        for (int j = lane; j < n; j += 32) {
            const double b = Br[j];
            const double* xj = x + (size_t)j * MV<C>::S;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fma(b, xj[c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = warp_sum(acc[c]);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) { double* yy = y + (size_t)row * MV<C>::S + c; *yy = add ? (*yy + acc[c]) : acc[c]; }
        }
    }
    __syncthreads();
}

// `steps` damped-Jacobi steps with the kernel correction on C right-hand sides at once (the
// multi-column twin of blk_smooth); r is the unit block (col0) when runit, else zero is never used.
template <int C>
__device__ double* smooth_multi(const LevelDev& L, int col0, double* ecur, double* ealt, bool e_zero, int steps, int isnsp,
                                const double (&sum_r)[C], double* red, int& flip) {
    const int sub = threadIdx.x % kBT;
    if (steps == 0) {
        if (e_zero) { for (int i = threadIdx.x; i < L.N * MV<C>::S; i += kDT) ecur[i] = 0.0; __syncthreads(); }
        return ecur;
    }
    double dotAe[C];
#pragma unroll
    for (int c = 0; c < C; ++c) dotAe[c] = 0.0;
    if (isnsp && !e_zero) {
        for (int i = threadIdx.x; i < L.N; i += kDT) {
            const double axi = L.Axi[i];
#pragma unroll
            for (int c = 0; c < C; ++c) dotAe[c] = fma(axi, ecur[(size_t)i * MV<C>::S + c], dotAe[c]);
        }
        block_sumC<C>(dotAe, red, flip);
    }
    for (int it = 0; it < steps; ++it) {
        double coef[C], part[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { coef[c] = isnsp ? (sum_r[c] - dotAe[c]) / L.xx : 0.0; part[c] = 0.0; }
        for (int base = 0; base < L.N; base += kDT / kBT) {
            const int row = base + threadIdx.x / kBT;
            const bool valid = row < L.N;
            double d[C];
            if (!e_zero) row_dot_multi<C>(L.ap, L.ai, L.av, ecur, row, sub, valid, d);
            else {
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = 0.0;
            }
            if (valid && sub == 0) {
                const double axi = L.Axi[row], di = L.dinv[row];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double ri = (row == col0 + c) ? 1.0 : 0.0;
                    const double gi = ri - d[c];
                    const double en = (e_zero ? 0.0 : ecur[(size_t)row * MV<C>::S + c]) + coef[c] + di * (gi - axi * coef[c]);
                    ealt[(size_t)row * MV<C>::S + c] = en;
                    part[c] = fma(axi, en, part[c]);
                }
            }
        }
        block_sumC<C>(part, red, flip);                   // the barrier inside also publishes ealt
#pragma unroll
        for (int c = 0; c < C; ++c) dotAe[c] = part[c];
        double* t = ecur; ecur = ealt; ealt = t;
        e_zero = false;
    }
    return ecur;
}

// B_k columns [C*blockIdx.x, +C): one level of the cycle applied to unit vectors
template <int C>
__global__ void __launch_bounds__(kDT) dense_build_kernel(LevelDev L, LevelDev Lc, const double* __restrict__ Bc, int smoth,
                                                          int isnsp, int twice, double* __restrict__ B) {
    extern __shared__ __align__(16) double dsm_d[];
    __shared__ double red[2 * C * 32];
    const int N = L.N, Nc = Lc.N;
    constexpr int CS = MV<C>::S;
    double* e0 = dsm_d; double* e1 = e0 + (size_t)N * CS; double* g = e1 + (size_t)N * CS;
    double* rc = g + (size_t)N * CS; double* ec = rc + (size_t)Nc * CS; double* dc = ec + (size_t)Nc * CS;
    const int col0 = blockIdx.x * C;
    int flip = 0;
    double sum_r[C];
#pragma unroll
    for (int c = 0; c < C; ++c) sum_r[c] = (col0 + c < N) ? 1.0 : 0.0;
    double* ecur = smooth_multi<C>(L, col0, e0, e1, true, smoth, isnsp, sum_r, red, flip);     // MG_Wcycle.m:15-24
    double* ealt = (ecur == e0) ? e1 : e0;
    {   // g = r - A e                                                                            :26
        const int sub = threadIdx.x % kBT;
        for (int base = 0; base < N; base += kDT / kBT) {
            const int row = base + threadIdx.x / kBT;
            const bool valid = row < N;
            double d[C];
            row_dot_multi<C>(L.ap, L.ai, L.av, ecur, row, sub, valid, d);
            if (valid && sub == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) g[(size_t)row * MV<C>::S + c] = ((row == col0 + c) ? 1.0 : 0.0) - d[c];
            }
        }
        __syncthreads();
    }
    spmm_smem<C>(Nc, Lc.tp, Lc.ti, Lc.tv, g, rc, 0);                                          // rc = Pro' g
    dense_smem<C>(Nc, Bc, rc, ec, false);                                                     // :28
    if (twice) {                                                                              // :30
        for (int i = threadIdx.x; i < Nc * CS; i += kDT) dc[i] = rc[i];
        __syncthreads();
        spmm_smem<C>(Nc, Lc.ap, Lc.ai, Lc.av, ec, dc, 2);                                     // dc = rc - A_c ec
        dense_smem<C>(Nc, Bc, dc, ec, true);
    }
    spmm_smem<C>(N, Lc.pp, Lc.pi, Lc.pv, ec, ecur, 1);                                        // :32
    ecur = smooth_multi<C>(L, col0, ecur, ealt, false, smoth, isnsp, sum_r, red, flip);       // :34-42
    for (int i = threadIdx.x; i < N * C; i += kDT) {
        const int row = i / C, c = i % C;
        if (col0 + c < N) B[(size_t)row * N + col0 + c] = ecur[(size_t)row * CS + c];
    }
}

// coarsest level: column i of B = PCG(A, e_i) with the defaults of PCG.m:18-23 (one warp, N <= 32)
__device__ double warp_pcg_lane(const LevelDev& L, double r) {
    const int lane = threadIdx.x & 31;
    const bool valid = lane < L.N;
    const int e0 = valid ? L.ap[lane] : 0, e1 = valid ? L.ap[lane + 1] : 0;
    int maxlen = e1 - e0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    const double diag = valid ? level_diag(L, lane) : 1.0;
    double p = r / diag, xi = 0.0;
    double delta_new = warp_sum(r * p);
    const double delta_0 = delta_new, tol2 = 1e-11 * 1e-11;
    int it = 0;
    while (it < 10000 && delta_new > tol2 * delta_0) {
        const double delta_old = delta_new;
        double q = 0.0;
        for (int t = 0; t < maxlen; ++t) {
            const int e = e0 + t;
            const bool has = e < e1;
            const int j = has ? L.ai[e] : 0;
            const double a = has ? L.av[e] : 0.0;
            const double pj = __shfl_sync(0xffffffffu, p, j);
            q = fma(a, pj, q);
        }
        const double alpha = delta_old / warp_sum(q * p);
        xi += alpha * p;
        r -= alpha * q;
        const double w = r / diag;
        delta_new = warp_sum(r * w);
        p = w + (delta_new / delta_old) * p;
        ++it;
    }
    return xi;
}

__global__ void __launch_bounds__(32) dense_coarsest_warp_kernel(LevelDev L, double* __restrict__ B) {
    const int col = blockIdx.x, lane = threadIdx.x;
    const double xi = warp_pcg_lane(L, (lane == col) ? 1.0 : 0.0);
    if (lane < L.N) B[(size_t)lane * L.N + col] = xi;
}

// coarsest level with more than 32 unknowns: block-wide PCG per column, vectors in shared memory
__global__ void __launch_bounds__(256) dense_coarsest_block_kernel(LevelDev L, double* __restrict__ B) {
    extern __shared__ __align__(16) double dsm_d[];
    __shared__ double red[32];
    const int n = L.N, col = blockIdx.x;
    double* x = dsm_d; double* r = x + n; double* p = r + n; double* q = p + n;
    double dn = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double ri = (i == col) ? 1.0 : 0.0, pi = ri / level_diag(L, i);
        r[i] = ri; p[i] = pi; x[i] = 0.0; dn = fma(ri, pi, dn);
    }
    double delta_new = block_sum(dn, red);
    const double delta_0 = delta_new, tol2 = 1e-11 * 1e-11;
    int it = 0;
    while (it < 10000 && delta_new > tol2 * delta_0) {
        const double delta_old = delta_new;
        double qp = 0.0;
        for (int i = threadIdx.x; i < n; i += 256) {
            double s = 0.0;
            for (int e = L.ap[i]; e < L.ap[i + 1]; ++e) s = fma(L.av[e], p[L.ai[e]], s);
            q[i] = s; qp = fma(s, p[i], qp);
        }
        qp = block_sum(qp, red);
        const double alpha = delta_old / qp;
        double dnew = 0.0;
        for (int i = threadIdx.x; i < n; i += 256) {
            x[i] += alpha * p[i];
            const double ri = r[i] - alpha * q[i];
            r[i] = ri; q[i] = ri / level_diag(L, i);
            dnew = fma(ri, q[i], dnew);
        }
        delta_new = block_sum(dnew, red);
        const double beta = delta_new / delta_old;
        for (int i = threadIdx.x; i < n; i += 256) p[i] = q[i] + beta * p[i];
        __syncthreads();
        ++it;
    }
    for (int i = threadIdx.x; i < n; i += 256) B[(size_t)i * n + col] = x[i];
}

// y (+)= B x, dense row-major n x n: one warp per row
__global__ void __launch_bounds__(256) dense_apply_kernel(int n, const double* __restrict__ B, const double* __restrict__ x,
                                                          double* __restrict__ y, int add) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const double* Br = B + (size_t)row * n;
    double s0 = 0.0, s1 = 0.0;
    int j = lane;
    for (; j + 32 < n; j += 64) { s0 = fma(Br[j], x[j], s0); s1 = fma(Br[j + 32], x[j + 32], s1); }
    if (j < n) s0 = fma(Br[j], x[j], s0);
    const double s = warp_sum(s0 + s1);
    if (lane == 0) y[row] = add ? (y[row] + s) : s;
}

// ------------------------------------------------------------------ cluster cycle kernel
//
// The single CTA above is bound by one SM's L2 bandwidth and issue rate.  For the mid-size
// levels (N <= 4096) a thread-block CLUSTER of 8 CTAs (8 SMs, hardware cluster barrier, DSMEM
// for the reductions) walks the sub-hierarchy instead: every CTA owns a contiguous row slice of
// every level and keeps that slice of the matrix in its shared memory when it fits; vectors live
// in global memory (L2) and are read with ld.global.cg after each cluster barrier.

constexpr int kCS = kClusterSize;              // CTAs per cluster (portable maximum)
constexpr int kCT = 1024;                      // threads per CTA
constexpr int kCL = kClusterLevels;            // levels a cluster kernel can hold

struct CLevel {                                // what a CTA uses for one level
    int r0, r1;                                // my rows
    const int* rp; const int* ci; const double* cv; int rbase;   // row pointers indexed by (row - rbase)
};

__device__ __forceinline__ double row_dot_cg(const int* __restrict__ rp, const int* __restrict__ ci,
                                             const double* __restrict__ cv, const double* __restrict__ x,
                                             int lrow, int sub, bool valid) {
    double s = 0.0;
    if (valid) {
        const int e1 = rp[lrow + 1];
        int e = rp[lrow] + sub;
        for (; e + 3 * kBT < e1; e += 4 * kBT) {
            const int i0 = ci[e], i1 = ci[e + kBT], i2 = ci[e + 2 * kBT], i3 = ci[e + 3 * kBT];
            const double v0 = cv[e], v1 = cv[e + kBT], v2 = cv[e + 2 * kBT], v3 = cv[e + 3 * kBT];
            const double x0 = __ldcg(x + i0), x1 = __ldcg(x + i1), x2 = __ldcg(x + i2), x3 = __ldcg(x + i3);
            s = fma(v0, x0, s); s = fma(v1, x1, s); s = fma(v2, x2, s); s = fma(v3, x3, s);
        }
        for (; e < e1; e += kBT) s = fma(cv[e], __ldcg(x + ci[e]), s);
    }
#pragma unroll
    for (int o = kBT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

struct ClusterRed { double blk[2][32]; double cl[2][kCS]; };

// cluster-wide sum, identical in every thread; includes one cluster barrier (release/acquire)
__device__ __forceinline__ double cl_sum(cg::cluster_group& cluster, double v, ClusterRed* red, int& flip, int rank) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = flip & 1;
    v = warp_sum(v);
    if (lane == 0) red->blk[b][w] = v;
    __syncthreads();
    if (w == 0) {
        double t = warp_sum(red->blk[b][lane]);
        if (lane < kCS) {
            double* remote = cluster.map_shared_rank(&red->cl[b][rank], lane);
            *remote = t;
        }
    }
    cluster.sync();
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kCS; ++r) s += red->cl[b][r];
    ++flip;
    return s;
}

// Jacobi smoothing with the kernel correction on my row slice; returns the buffer holding e
__device__ double* cl_smooth(cg::cluster_group& cluster, const LevelDev& L, const CLevel& C, const double* r, double* ecur,
                             double* ealt, bool e_zero, int steps, int isnsp, double sum_r, ClusterRed* red, int& flip, int rank) {
    const int sub = threadIdx.x % kBT;
    if (steps == 0) {
        if (e_zero) { for (int i = C.r0 + threadIdx.x; i < C.r1; i += kCT) ecur[i] = 0.0; cluster.sync(); }
        return ecur;
    }
    double dotAe = 0.0;
    if (isnsp && !e_zero) {
        double s = 0.0;
        for (int i = C.r0 + threadIdx.x; i < C.r1; i += kCT) s = fma(L.Axi[i], __ldcg(ecur + i), s);
        dotAe = cl_sum(cluster, s, red, flip, rank);
    }
    for (int it = 0; it < steps; ++it) {
        const double coef = isnsp ? (sum_r - dotAe) / L.xx : 0.0;
        double part = 0.0;
        for (int base = C.r0; base < C.r1; base += kCT / kBT) {
            const int row = base + threadIdx.x / kBT;
            const bool valid = row < C.r1;
            double d = 0.0;
            if (!e_zero) d = row_dot_cg(C.rp, C.ci, C.cv, ecur, row - C.rbase, sub, valid);
            if (valid && sub == 0) {
                const double axi = L.Axi[row];
                const double gi = __ldcg(r + row) - d;
                const double en = (e_zero ? 0.0 : __ldcg(ecur + row)) + coef + L.dinv[row] * (gi - axi * coef);
                ealt[row] = en;
                part = fma(axi, en, part);
            }
        }
        dotAe = cl_sum(cluster, part, red, flip, rank);   // its barrier publishes ealt cluster-wide
        double* t = ecur; ecur = ealt; ealt = t;
        e_zero = false;
    }
    return ecur;
}

// y[rows of my slice of an nrows-row matrix] (+)= M*x
__device__ void cl_spmv(cg::cluster_group& cluster, int nrows, int rank, const int* ptr, const int* idx, const double* val,
                        const double* x, double* y, bool add) {
    const int rpc = (nrows + kCS - 1) / kCS;
    const int r0 = min(nrows, rank * rpc), r1 = min(nrows, r0 + rpc);
    const int sub = threadIdx.x % kBT;
    for (int base = r0; base < r1; base += kCT / kBT) {
        const int row = base + threadIdx.x / kBT;
        const bool valid = row < r1;
        const double d = row_dot_cg(ptr, idx, val, x, row, sub, valid);
        if (valid && sub == 0) y[row] = add ? (__ldcg(y + row) + d) : d;
    }
    cluster.sync();
}

__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kCT)
cluster_cycle_kernel(const LevelDev* __restrict__ levels, int k0, int J, int smoth, int isnsp, int wcycle, int e0_zero,
                     const ClusterPlan plan) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ ClusterRed red;
    __shared__ LevelDev sl[kCL];
    __shared__ CLevel cls[kCL];
    const int rank = (int)cluster.block_rank();
    const int nl = J - k0;
    for (int t = threadIdx.x; t < nl; t += kCT) sl[t] = levels[k0 + t];
    __syncthreads();
    // ---- stage my matrix slices
    for (int t = 0; t < nl; ++t) {
        const LevelDev& L = sl[t];
        const int rpc = plan.rpc[t];
        const int r0 = min(L.N, rank * rpc), r1 = min(L.N, r0 + rpc);
        if (plan.staged[t]) {
            int* sp = reinterpret_cast<int*>(dsm + plan.smem_off[t]);
            double* sv = reinterpret_cast<double*>(dsm + plan.smem_off[t] + (((rpc + 1) * 4 + 15) / 16) * 16);
            int* si = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(sv) + (size_t)plan.cap[t] * 8);
            const int base = L.ap[r0];
            for (int i = threadIdx.x; i <= r1 - r0; i += kCT) sp[i] = L.ap[r0 + i] - base;
            const int cnt = L.ap[r1] - base;
            for (int e = threadIdx.x; e < cnt; e += kCT) { sv[e] = L.av[base + e]; si[e] = L.ai[base + e]; }
            if (threadIdx.x == 0) { cls[t].r0 = r0; cls[t].r1 = r1; cls[t].rp = sp; cls[t].ci = si; cls[t].cv = sv; cls[t].rbase = r0; }
        } else if (threadIdx.x == 0) {
            cls[t].r0 = r0; cls[t].r1 = r1; cls[t].rp = L.ap; cls[t].ci = L.ai; cls[t].cv = L.av; cls[t].rbase = 0;
        }
    }
    cluster.sync();
    int flip = 0;
    int phase[kCL]; bool zero[kCL]; double* ecur[kCL]; double* ealt[kCL]; double sum_r[kCL];
    for (int t = 0; t < kCL; ++t) { ecur[t] = nullptr; ealt[t] = nullptr; sum_r[t] = 0.0; phase[t] = 0; zero[t] = true; }
    for (int t = 0; t < nl; ++t) { ecur[t] = sl[t].e; ealt[t] = sl[t].pcg; }
    int k = 0;
    zero[0] = e0_zero != 0;
    while (true) {
        const LevelDev& L = sl[k];
        const CLevel& C = cls[k];
        if (k == nl - 1) {                                // coarsest (N <= 32, checked by the host): warp PCG on CTA 0
            if (rank == 0 && threadIdx.x < 32) warp_pcg(L, L.r, L.e);
            cluster.sync();
            ecur[k] = L.e;
            if (k == 0) break;
            --k; continue;
        }
        if (phase[k] == 0) {
            if (isnsp) {
                double s = 0.0;
                for (int i = C.r0 + threadIdx.x; i < C.r1; i += kCT) s += __ldcg(L.r + i);
                sum_r[k] = cl_sum(cluster, s, &red, flip, rank);
            }
            double* en = cl_smooth(cluster, L, C, L.r, ecur[k], ealt[k], zero[k], smoth, isnsp, sum_r[k], &red, flip, rank);
            if (en != ecur[k]) { ealt[k] = ecur[k]; ecur[k] = en; }
            {   // g = r - A e on my rows, then the restriction (needs every CTA's g)
                const int sub = threadIdx.x % kBT;
                for (int base = C.r0; base < C.r1; base += kCT / kBT) {
                    const int row = base + threadIdx.x / kBT;
                    const bool valid = row < C.r1;
                    const double d = row_dot_cg(C.rp, C.ci, C.cv, ecur[k], row - C.rbase, sub, valid);
                    if (valid && sub == 0) L.g[row] = __ldcg(L.r + row) - d;
                }
                cluster.sync();
            }
            const LevelDev& Lc = sl[k + 1];
            cl_spmv(cluster, Lc.N, rank, Lc.tp, Lc.ti, Lc.tv, L.g, Lc.r, false);
            phase[k] = 1; phase[k + 1] = 0; zero[k + 1] = true; ecur[k + 1] = Lc.e; ealt[k + 1] = Lc.pcg; ++k; continue;
        }
        if (phase[k] == 1 && wcycle && (k + 1 != nl - 1)) {
            phase[k] = 2; phase[k + 1] = 0; zero[k + 1] = false; ++k; continue;
        }
        {
            const LevelDev& Lc = sl[k + 1];
            cl_spmv(cluster, L.N, rank, Lc.pp, Lc.pi, Lc.pv, ecur[k + 1], ecur[k], true);
            double* en = cl_smooth(cluster, L, C, L.r, ecur[k], ealt[k], false, smoth, isnsp, sum_r[k], &red, flip, rank);
            if (en != ecur[k]) { ealt[k] = ecur[k]; ecur[k] = en; }
        }
        if (k == 0) break;
        --k;
    }
    if (ecur[0] != sl[0].e) {
        const CLevel& C = cls[0];
        for (int i = C.r0 + threadIdx.x; i < C.r1; i += kCT) sl[0].e[i] = __ldcg(ecur[0] + i);
    }
}

__global__ void slice_bounds_kernel(const LevelDev* __restrict__ levels, int J, int* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= J * (kCS + 1)) return;
    const int k = t / (kCS + 1), r = t % (kCS + 1);
    const LevelDev& L = levels[k];
    const int rpc = (L.N + kCS - 1) / kCS;
    const int row = min(L.N, r * rpc);
    out[t] = L.ap[row];
}

// ------------------------------------------------------------------ general PCG (persistent, cooperative)

struct PcgArgs {
    int n; const int* ptr; const int* idx; const double* val;
    const double* rhs; const double* guess;
    double* d; double* r; double* p; double* q; double* w; double* a;   // a: extra vector for bi-SSOR
    const double* diag;
    int precd, nf, maxit; double tol2;
    double* part;            // [3][gridDim]
    double* resk;            // maxit
    int* it_out; double* scal_out;   // scal_out[0] = delta_new, [1] = delta_0
    // precd 3 / 4: w = Uf \ (mid .* (Lf \ r)) with level-scheduled sparse triangular solves
    const int* lp; const int* li; const double* lv;     // Lf, CSR, columns ascending, diagonal LAST in its row
    const int* up; const int* ui; const double* uv;     // Uf, CSR, columns ascending, diagonal FIRST in its row
    const double* mid;                                  // optional diagonal scaling between the solves
    const int* lrows; const int* llev; int nlev_l;      // rows of Lf grouped by dependency level, level pointers
    const int* urows; const int* ulev; int nlev_u;
};

__device__ __forceinline__ double grid_sum(cg::grid_group& grid, double v, double* part, double* red) {
    v = block_sum(v, red);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    grid.sync();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += part[i];
    s = block_sum(s, red);
    return s;
}

// w = M^{-1} r for precd 1, 2, 5 (PCG.m:90-105); returns this thread's share of r'w
__device__ double pcg_precond(cg::grid_group& grid, const PcgArgs& a, int gtid, int gsize) {
    double dn = 0.0;
    if (a.precd == 1) {
        for (int i = gtid; i < a.n; i += gsize) { const double ri = a.r[i]; a.w[i] = ri; dn = fma(ri, ri, dn); }
    } else if (a.precd == 2) {
        for (int i = gtid; i < a.n; i += gsize) { const double ri = a.r[i]; const double wi = ri / a.diag[i]; a.w[i] = wi; dn = fma(ri, wi, dn); }
    } else if (a.precd == 3 || a.precd == 4) {
        // SSOR (PCG.m:96-99) / incomplete Cholesky (:100-101): forward solve into a.a, backward solve into a.w.
        // Rows of one dependency level are independent: a warp per row, a grid barrier per level.
        const int lane = threadIdx.x & 31, gwarp = gtid >> 5, nwarps = gsize >> 5;
        for (int lev = 0; lev < a.nlev_l; ++lev) {
            for (int t = a.llev[lev] + gwarp; t < a.llev[lev + 1]; t += nwarps) {
                const int row = a.lrows[t];
                const int e0 = a.lp[row], e1 = a.lp[row + 1] - 1;           // e1: the diagonal entry
                double s = 0.0;
                for (int e = e0 + lane; e < e1; e += 32) s = fma(a.lv[e], a.a[a.li[e]], s);
                s = warp_sum(s);
                if (lane == 0) a.a[row] = (a.r[row] - s) / a.lv[e1];
            }
            grid.sync();
        }
        for (int lev = 0; lev < a.nlev_u; ++lev) {
            for (int t = a.ulev[lev] + gwarp; t < a.ulev[lev + 1]; t += nwarps) {
                const int row = a.urows[t];
                const int e0 = a.up[row], e1 = a.up[row + 1];               // e0: the diagonal entry
                double s = 0.0;
                for (int e = e0 + 1 + lane; e < e1; e += 32) s = fma(a.uv[e], a.w[a.ui[e]], s);
                s = warp_sum(s);
                if (lane == 0) { const double rhs = a.mid ? a.mid[row] * a.a[row] : a.a[row]; a.w[row] = (rhs - s) / a.uv[e0]; }
            }
            grid.sync();
        }
        for (int i = gtid; i < a.n; i += gsize) dn = fma(a.r[i], a.w[i], dn);
    } else {
        // bi-SSOR, w = 1.5:  P r = w(2-w) [ aa - w invV U b ; b ],  aa = invV r_c,  b = invT (r_r - w U' aa)
        const double om = 1.5, sc = om * (2.0 - om);
        for (int i = gtid; i < a.nf; i += gsize) a.a[i] = a.r[i] / a.diag[i];
        grid.sync();
        for (int i = a.nf + gtid; i < a.n; i += gsize) {
            double s = 0.0;
            for (int e = a.ptr[i]; e < a.ptr[i + 1]; ++e) { const int j = a.idx[e]; if (j < a.nf) s = fma(a.val[e], a.a[j], s); }
            a.a[i] = (a.r[i] - om * s) / a.diag[i];
        }
        grid.sync();
        for (int i = gtid; i < a.n; i += gsize) {
            double wi;
            if (i < a.nf) {
                double s = 0.0;
                for (int e = a.ptr[i]; e < a.ptr[i + 1]; ++e) { const int j = a.idx[e]; if (j >= a.nf) s = fma(a.val[e], a.a[j], s); }
                wi = sc * (a.a[i] - om * s / a.diag[i]);
            } else wi = sc * a.a[i];
            a.w[i] = wi; dn = fma(a.r[i], wi, dn);
        }
    }
    return dn;
}

__global__ void __launch_bounds__(256) pcg_kernel(PcgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[32];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int gwarp = gtid >> 5, nwarps = gsize >> 5;
    double* part0 = a.part; double* part1 = a.part + gridDim.x; double* part2 = a.part + 2 * gridDim.x;
    // r = e - H*d0 ; d = d0                                   (PCG.m:68-70)
    for (int i = gtid; i < a.n; i += gsize) a.d[i] = a.guess ? a.guess[i] : 0.0;
    grid.sync();
    for (int row = gwarp; row < a.n; row += nwarps) {
        double s = 0.0;
        if (a.guess) for (int e = a.ptr[row] + lane; e < a.ptr[row + 1]; e += 32) s = fma(a.val[e], a.d[a.idx[e]], s);
        s = warp_sum(s);
        if (lane == 0) a.r[row] = a.rhs[row] - s;
    }
    grid.sync();
    double dn = pcg_precond(grid, a, gtid, gsize);
    double delta_new = grid_sum(grid, dn, part0, red);
    for (int i = gtid; i < a.n; i += gsize) a.p[i] = a.w[i];
    const double delta_0 = delta_new;
    int it = 0;
    grid.sync();
    while (it < a.maxit && delta_new > a.tol2 * delta_0) {           // PCG.m:76
        const double delta_old = delta_new;
        double qp = 0.0;
        for (int row = gwarp; row < a.n; row += nwarps) {            // q = H*p
            double s = 0.0;
            for (int e = a.ptr[row] + lane; e < a.ptr[row + 1]; e += 32) s = fma(a.val[e], a.p[a.idx[e]], s);
            s = warp_sum(s);
            if (lane == 0) { a.q[row] = s; qp = fma(s, a.p[row], qp); }
        }
        qp = grid_sum(grid, qp, part1, red);
        const double alpha = delta_old / qp;
        for (int i = gtid; i < a.n; i += gsize) { a.d[i] += alpha * a.p[i]; a.r[i] -= alpha * a.q[i]; }
        if (a.precd >= 3) grid.sync();
        dn = pcg_precond(grid, a, gtid, gsize);
        delta_new = grid_sum(grid, dn, (it & 1) ? part0 : part2, red);
        const double beta = delta_new / delta_old;
        for (int i = gtid; i < a.n; i += gsize) a.p[i] = a.w[i] + beta * a.p[i];
        ++it;
        if (gtid == 0 && a.resk) a.resk[it - 1] = sqrt(fabs(delta_new / delta_0));
        grid.sync();
    }
    if (gtid == 0) { *a.it_out = it; a.scal_out[0] = delta_new; a.scal_out[1] = delta_0; }
}

// ------------------------------------------------------------------ persistent solve kernel
//
// Class_AMG's whole solve loop (Class_AMG.m:89-107: r = b - A*x, x += cycle(r), stop on the
// relative residual / divergence guard) as ONE cooperative kernel: every dependent step of the
// cycle on the large levels (fused Jacobi step, block Gauss-Seidel residual + coupled update,
// restriction, prolongation, dense tail operator) is a grid-wide pass followed by a grid barrier
// (~1-2 us) instead of a kernel launch (~5 us of launch + ramp + tail), reductions ride on the
// same barrier, and the convergence test runs on the device.  Vectors that change during the
// kernel are read with ld.global.cg (L2), never through the non-coherent path.

#ifndef SSN_PT
#define SSN_PT 256
#endif
#ifndef SSN_PBPS
#define SSN_PBPS 2
#endif
constexpr int kPT = SSN_PT;                   // threads per block of the persistent kernel
constexpr int kPBlocksPerSM = SSN_PBPS;       // resident blocks per SM (the grid is num_sms * kPBlocksPerSM)
constexpr int kPLevels = 16;
constexpr int kGridXs = 8192;                  // doubles of shared memory per block of the grid-wide kernel for the copy of a gathered vector

struct PersistArgs {
    const LevelDev* levels; int J, kd, smoth, isnsp, wcycle;
    const double* b; double* x; double retol; int maxit;
    double* part;                 // [2][2*gridDim.x]
    double* relk; double* rho;    // maxit + 2 entries each
    int* it_out;                  // [0] = it, [1] = history length
    int tpr[kPLevels];            // lanes per row of A_k (k < kd)
    int tpr_p[kPLevels];          // lanes per row of Pro_k / Pro_k' (k <= kd)
    int tpr_gs;                   // lanes per row of the coupled block Gauss-Seidel update (0: tpr[0])
    int xs_ok[kPLevels];          // 1: A_k is dense enough that its passes read the gathered vector from a shared-memory copy
    size_t smem_budget;           // dynamic shared memory available for staged matrix slices
};

// A block's contiguous slice of rows of one CSR matrix: either staged in shared memory (row
// pointers rebased to the slice) or read in place from global memory.
struct PSlice {
    int r0, r1, rbase; const int* rp; const int* ci; const double* cv;
    int inter;     // 1: rows interleaved over the whole grid (in-place matrices only); 0: the block's contiguous slice
    __device__ __forceinline__ int first(int rows_per_pass, int blk) const { return inter ? blk * rows_per_pass : r0; }
    __device__ __forceinline__ int step(int rows_per_pass, int nblk) const { return inter ? nblk * rows_per_pass : rows_per_pass; }
};
struct PLevelS { PSlice A, Pu, Td; };      // A_t ; Pro_{t+1} (rows of level t) ; Pro_t' (rows of level t)
struct PState {                            // control state of the cycle, one entry per level (shared memory)
    int phase[kPLevels]; int zero[kPLevels]; double* ecur[kPLevels]; double* ealt[kPLevels]; double sum_r[kPLevels]; double dot_e[kPLevels];
};

template <class TM>
__device__ void stage_slice(const TM& G, PSlice* S, int N, const int* ptr, const int* idx, const double* val, unsigned char* smem,
                            size_t& used, size_t budget) {
    constexpr int kPT = TM::T;
    const int rpb = (N + G.nblk() - 1) / G.nblk();
    const int r0 = min(N, G.blk() * rpb), r1 = min(N, r0 + rpb);
    const int nrows = r1 - r0;
    const int base = (ptr && nrows > 0) ? ptr[r0] : 0;
    const int cnt = (ptr && nrows > 0) ? (ptr[r1] - base) : 0;
    const size_t bv = ((size_t)cnt * 8 + 15) / 16 * 16, bi = ((size_t)cnt * 4 + 15) / 16 * 16, bp = ((size_t)(nrows + 1) * 4 + 15) / 16 * 16;
    const bool fits = ptr != nullptr && nrows > 0 && used + bv + bi + bp <= budget;
    if (fits) {
        double* sv = reinterpret_cast<double*>(smem + used);
        int* si = reinterpret_cast<int*>(smem + used + bv);
        int* sp = reinterpret_cast<int*>(smem + used + bv + bi);
        for (int i = threadIdx.x; i <= nrows; i += kPT) sp[i] = ptr[r0 + i] - base;
        for (int e = threadIdx.x; e < cnt; e += kPT) { sv[e] = val[base + e]; si[e] = idx[base + e]; }
        if (threadIdx.x == 0) { S->r0 = r0; S->r1 = r1; S->rbase = r0; S->rp = sp; S->ci = si; S->cv = sv; S->inter = 0; }
        used += bv + bi + bp;
    } else if (threadIdx.x == 0) {
        const int inter = (budget == 0) ? 1 : 0;             // no staging at all: spread the rows over the grid
        S->r0 = inter ? 0 : r0; S->r1 = inter ? N : r1; S->rbase = 0; S->rp = ptr; S->ci = idx; S->cv = val; S->inter = inter;
    }
}

// SM: x is a shared-memory copy of the vector (plain loads) instead of the vector in global memory (ld.global.cg)
template <int TPR, bool SM = false>
__device__ __forceinline__ double row_dot_s(const PSlice& S, const double* x, int row, int sub, bool valid) {
    double s = 0.0;
    if (valid) {
        const int e1 = S.rp[row - S.rbase + 1];
        int e = S.rp[row - S.rbase] + sub;
        const int* ci = S.ci; const double* cv = S.cv;
        // dense rows (hundreds of entries: early-phase systems, the coarse levels of partial OT): eight index / value loads of a
        // lane in flight per trip -- a lane's trips are dependent L2 round trips and the longest row of a block sets the time of
        // the pass (Class 2 bench state, 860 entries per row: 9 trips per row with batches of four and a one-by-one tail).
        // The products are added in entry order whatever the batching: same bits.
        if constexpr (SM)                                   // (the staged dense levels of the grid-wide kernel only: 128 registers there)
        for (; e + 7 * TPR < e1; e += 8 * TPR) {
            int iv[8]; double vv[8], xx[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { iv[u] = ci[e + u * TPR]; vv[u] = cv[e + u * TPR]; }
#pragma unroll
            for (int u = 0; u < 8; ++u) xx[u] = SM ? x[iv[u]] : __ldcg(x + iv[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) s = fma(vv[u], xx[u], s);
        }
        for (; e + 3 * TPR < e1; e += 4 * TPR) {
            const int i0 = ci[e], i1 = ci[e + TPR], i2 = ci[e + 2 * TPR], i3 = ci[e + 3 * TPR];
            const double v0 = cv[e], v1 = cv[e + TPR], v2 = cv[e + 2 * TPR], v3 = cv[e + 3 * TPR];
            const double x0 = SM ? x[i0] : __ldcg(x + i0), x1 = SM ? x[i1] : __ldcg(x + i1), x2 = SM ? x[i2] : __ldcg(x + i2), x3 = SM ? x[i3] : __ldcg(x + i3);
            s = fma(v0, x0, s); s = fma(v1, x1, s); s = fma(v2, x2, s); s = fma(v3, x3, s);
        }
        if constexpr (!SM) { for (; e < e1; e += TPR) s = fma(cv[e], __ldcg(x + ci[e]), s); }
        else if (e < e1) {                                  // the last one to three entries together (they were one round trip each)
            int iv[3]; double vv[3], xx[3]; bool has[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) { has[u] = e + u * TPR < e1; iv[u] = has[u] ? ci[e + u * TPR] : 0; vv[u] = has[u] ? cv[e + u * TPR] : 0.0; }
#pragma unroll
            for (int u = 0; u < 3; ++u) xx[u] = has[u] ? (SM ? x[iv[u]] : __ldcg(x + iv[u])) : 0.0;
#pragma unroll
            for (int u = 0; u < 3; ++u) if (has[u]) s = fma(vv[u], xx[u], s);
        }
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

template <class F>
__device__ __forceinline__ void with_tpr(int tpr, F&& f) {
    switch (tpr) {
        case 1: f(std::integral_constant<int, 1>()); break;
        case 2: f(std::integral_constant<int, 2>()); break;
        case 4: f(std::integral_constant<int, 4>()); break;
        case 8: f(std::integral_constant<int, 8>()); break;
        case 16: f(std::integral_constant<int, 16>()); break;
        default: f(std::integral_constant<int, 32>()); break;
    }
}

// Grid-wide barrier for the persistent kernel (one block per SM, all resident: cooperative launch).
// bar[0] = arrival counter, bar[32] = generation (separate 128-byte lines so that the pollers do not
// slow the arrivals down).  Same guarantees as cg::grid_group::sync() -- every global write made
// before the barrier is visible to every thread after it -- at roughly half the latency.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned* gen = bar + 32;
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 32, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPT, kPBlocksPerSM) barrier_bench_kernel(unsigned* bar, int iters, int which, long long* cycles_out) {
    cg::grid_group grid = cg::this_grid();
    grid.sync();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (which == 0) grid.sync(); else grid_barrier(bar, gridDim.x); }
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles_out[0] = clock64() - t0;
}

// The "team" that runs the persistent solve body: either the whole cooperative grid (GridTeam, one grid barrier per
// dependent pass) or ONE thread-block cluster (ClusterTeam: hardware cluster barrier, reductions through distributed
// shared memory).  Same interface: T threads per block, blk() / nblk(), sync(), sum2().
struct GridTeam {
    static constexpr int T = kPT;
    static constexpr int U = 1;               // rows per thread in flight (row_dots)
    cg::grid_group grid; double* part; double* red; int flip;
    double* xs = nullptr;                     // shared-memory copy of a gathered vector (dense levels, the dense tail input)
    int xs_cap = 0;
    // vectors that change during the kernel are read from L2, never through a stale L1 line
    static __device__ __forceinline__ double ld(const double* p) { return __ldcg(p); }
    __device__ __forceinline__ int blk() const { return (int)blockIdx.x; }
    __device__ __forceinline__ int nblk() const { return (int)gridDim.x; }
    __device__ __forceinline__ void sync() { grid.sync(); }
    // grid-wide sums of two per-thread values; one grid barrier; results identical in every thread
    __device__ __forceinline__ void sum2(double& a, double& b) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        a = warp_sum(a); b = warp_sum(b);
        double* sm = red + (flip & 1) * 64;
        if (lane == 0) { sm[w] = a; sm[32 + w] = b; }
        __syncthreads();
        double* p = part + (size_t)(flip & 1) * 2 * gridDim.x;
        if (w == 0) {
            double ta = (lane < kPT / 32) ? sm[lane] : 0.0, tb = (lane < kPT / 32) ? sm[32 + lane] : 0.0;
            ta = warp_sum(ta); tb = warp_sum(tb);
            if (lane == 0) { p[2 * blockIdx.x] = ta; p[2 * blockIdx.x + 1] = tb; }
        }
        grid.sync();
        {   // every warp fetches its share of the block partials at once (one L2 round trip), fixed order
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += kPT) { s0 += __ldcg(p + 2 * i); s1 += __ldcg(p + 2 * i + 1); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (lane == 0) { sm[w] = s0; sm[32 + w] = s1; }
        }
        __syncthreads();
        {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < kPT / 32; ++i) { s0 += sm[i]; s1 += sm[32 + i]; }
            a = s0; b = s1;
        }
        ++flip;
    }
};

#ifndef SSN_CTT
#define SSN_CTT 1024
#endif
#ifndef SSN_CU
#define SSN_CU 4
#endif
constexpr int kCTT = SSN_CTT;                 // threads per CTA of the cluster-resident solve kernel
constexpr int kCMax = 16;                     // largest cluster (non-portable size, opt-in)

__device__ __forceinline__ void cluster_barrier() {
    // release / acquire at cluster scope: every global and shared::cluster write made before the barrier by any
    // CTA of the cluster is visible to every thread of the cluster after it
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct ClusterTeam {
    static constexpr int T = kCTT;
    static constexpr int U = SSN_CU;          // rows per thread in flight: 16 SMs carry the whole level, so a pass is bound by
                                              // the L2 round trips a thread can overlap, not by the barrier
    double* red;                              // [2][64] warp partials of this CTA
    double* slots;                            // [2][2*kCMax]: slot r of every CTA is written by CTA r through DSMEM
    int rank, ncta, flip;
    double* xs;                               // 2048 doubles: the input vector of the dense tail operator
    int xs_cap = 2048;
    // Plain (L1-cached) loads: every pass ends in a cluster barrier with release / acquire semantics, for which the
    // compiler emits CCTL.IVALL -- the L1 is invalidated at every barrier, so a line can only have been filled after
    // the last write to it by another CTA, and the gathers of a row slice, which mostly hit the slice's own
    // neighbourhood, are served by the L1 instead of one L2 sector request each.
    static __device__ __forceinline__ double ld(const double* p) { return *p; }      // ld.global (not .volatile / .cg / .nc)
    __device__ __forceinline__ int blk() const { return rank; }
    __device__ __forceinline__ int nblk() const { return ncta; }
    __device__ __forceinline__ void sync() { cluster_barrier(); }
    // cluster-wide sums of two per-thread values; ONE cluster barrier, no trip through L2; fixed order
    __device__ __forceinline__ void sum2(double& a, double& b) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        a = warp_sum(a); b = warp_sum(b);
        double* sm = red + (flip & 1) * 64;
        if (lane == 0) { sm[w] = a; sm[32 + w] = b; }
        __syncthreads();
        double* sl = slots + (flip & 1) * (2 * kCMax);
        if (w == 0) {
            double ta = (lane < T / 32) ? sm[lane] : 0.0, tb = (lane < T / 32) ? sm[32 + lane] : 0.0;
            ta = warp_sum(ta); tb = warp_sum(tb);
            if (lane < ncta) {
                cg::cluster_group cl = cg::this_cluster();
                double* remote = cl.map_shared_rank(sl + 2 * rank, lane);
                remote[0] = ta; remote[1] = tb;
            }
        }
        cluster_barrier();
        double s0 = 0.0, s1 = 0.0;
        for (int r = 0; r < ncta; ++r) { s0 += sl[2 * r]; s1 += sl[2 * r + 1]; }
        a = s0; b = s1;
        ++flip;
    }
};

// d[u] = A(row_u,:)*x for the U rows row0 + u*RP of a thread (TPR lanes per row), all U rows in flight at once: every
// round issues one index/value load and one gather per row before any of them is consumed, so a thread has U
// independent L2 round trips outstanding instead of one.  U == 1 keeps the 4-deep batching inside the row.
template <int TPR, int U, class TM, bool SM = false>
__device__ __forceinline__ void row_dots(const PSlice& S, const double* x, int row0, int RP, int sub, double (&d)[U]) {
    if constexpr (U == 1) {
        d[0] = row_dot_s<TPR, SM>(S, x, row0, sub, row0 < S.r1);
    } else {
        int e[U], e1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int row = row0 + u * RP;
            const bool valid = row < S.r1;
            e[u] = valid ? S.rp[row - S.rbase] + sub : 0; e1[u] = valid ? S.rp[row - S.rbase + 1] : 0; d[u] = 0.0;
        }
        bool more;
        do {
            int i[U]; double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const bool has = e[u] < e1[u]; i[u] = has ? S.ci[e[u]] : -1; v[u] = has ? S.cv[e[u]] : 0.0; }
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = (i[u] >= 0) ? (SM ? x[i[u]] : TM::ld(x + i[u])) : 0.0;
            more = false;
#pragma unroll
            for (int u = 0; u < U; ++u) { d[u] = fma(v[u], xv[u], d[u]); e[u] += TPR; more |= (e[u] < e1[u]); }
        } while (more);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
    }
}

// A level whose matrix is dense (hundreds of entries per row: early-phase SsN systems, the coarse levels of partial OT)
// re-reads every element of the gathered vector many times per pass: each block copies the n <= xs_cap doubles into its
// shared memory once and gathers from there (its own L2 traffic drops from 20 to 12 bytes per entry and the gathers stop
// being dependent L2 round trips).  Returns the copy, or null when the level does not qualify.
template <class TM>
__device__ __forceinline__ const double* p_stage_x(TM& G, const double* x, int n, bool on) {
    if (!on || x == nullptr || G.xs == nullptr || n > G.xs_cap) return nullptr;
    __syncthreads();                                        // nobody still reads the previous copy
    for (int i = threadIdx.x; i < n; i += TM::T) G.xs[i] = TM::ld(x + i);
    __syncthreads();
    return G.xs;
}

// g = r - A e on the block's rows (e == nullptr: g = r); sum g and sum g^2 when want (else just the barrier)
template <class TM>
__device__ void p_resid(TM& G, const PSlice& S, int tpr, const double* r, const double* e, double* g, bool want,
                        double& sg, double& sg2, int n_stage = 0) {
    sg = 0.0; sg2 = 0.0;
    const double* xs = p_stage_x(G, e, n_stage, n_stage > 0);
    with_tpr(tpr, [&](auto T) {
        constexpr int TPR = decltype(T)::value;
        constexpr int U = TM::U, RP = TM::T / TPR;
        const int sub = threadIdx.x % TPR;
        for (int base = S.first(U * RP, G.blk()); base < S.r1; base += S.step(U * RP, G.nblk())) {
            const int row0 = base + threadIdx.x / TPR;
            double d[U];
            if (xs != nullptr) row_dots<TPR, U, TM, true>(S, xs, row0, RP, sub, d);
            else if (e != nullptr) row_dots<TPR, U, TM>(S, e, row0, RP, sub, d);
            else {
#pragma unroll
                for (int u = 0; u < U; ++u) d[u] = 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = row0 + u * RP;
                if (row < S.r1 && sub == 0) { const double gi = TM::ld(r + row) - d[u]; g[row] = gi; sg += gi; sg2 = fma(gi, gi, sg2); }
            }
        }
    });
    if (want) G.sum2(sg, sg2); else G.sync();
}

// block Gauss-Seidel coupled update (Class_AMG.m:56-59 / MG_Wcycle.m:19,37): e (+)= coef + R(g - Axi*coef)
template <class TM>
__device__ void p_gs_apply(TM& G, const LevelDev& L, const PSlice& S, int tpr, const double* g, double* e, double coef, int post,
                           bool e_zero) {
    with_tpr(tpr, [&](auto T) {
        constexpr int TPR = decltype(T)::value;
        const int sub = threadIdx.x % TPR;
        for (int base = S.first(TM::T / TPR, G.blk()); base < S.r1; base += S.step(TM::T / TPR, G.nblk())) {
            const int row = base + threadIdx.x / TPR;
            const bool valid = row < S.r1;
            double hi = 0.0, di = 0.0;
            if (valid) { hi = TM::ld(g + row) - L.Axi[row] * coef; di = L.dinv[row]; }
            double s = 0.0;
            const bool coupled = valid && (post ? (row < L.Nf) : (row >= L.Nf));
            if (coupled) {
                // 4 entries per batch: all index loads, then all gathers, then the arithmetic (one dependent L2 round
                // trip per batch instead of one per entry)
                const int e1 = S.rp[row - S.rbase + 1];
                int q = S.rp[row - S.rbase] + sub;
                for (; q + 3 * TPR < e1; q += 4 * TPR) {
                    int j[4]; double av[4], gj[4], dj[4], xj[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { j[u] = S.ci[q + u * TPR]; av[u] = S.cv[q + u * TPR]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) { gj[u] = TM::ld(g + j[u]); dj[u] = L.dinv[j[u]]; xj[u] = L.Axi[j[u]]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool other = post ? (j[u] >= L.Nf) : (j[u] < L.Nf);
                        if (other) s = fma(av[u], dj[u] * (gj[u] - xj[u] * coef), s);
                    }
                }
                if constexpr (TM::T != kPT) {               // (the 64-register cluster kernels keep the one-by-one tail)
                    for (; q < e1; q += TPR) {
                        const int j = S.ci[q];
                        const bool other = post ? (j >= L.Nf) : (j < L.Nf);
                        if (other) s = fma(S.cv[q], L.dinv[j] * (TM::ld(g + j) - L.Axi[j] * coef), s);
                    }
                } else if (q < e1) {                        // the last one to three entries together
                    int j[3]; double av[3], gj[3], dj[3], xj[3]; bool has[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) { has[u] = q + u * TPR < e1; j[u] = has[u] ? S.ci[q + u * TPR] : 0; av[u] = has[u] ? S.cv[q + u * TPR] : 0.0; }
#pragma unroll
                    for (int u = 0; u < 3; ++u) { gj[u] = has[u] ? TM::ld(g + j[u]) : 0.0; dj[u] = has[u] ? L.dinv[j[u]] : 0.0; xj[u] = has[u] ? L.Axi[j[u]] : 0.0; }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const bool other = has[u] && (post ? (j[u] >= L.Nf) : (j[u] < L.Nf));
                        if (other) s = fma(av[u], dj[u] * (gj[u] - xj[u] * coef), s);
                    }
                }
            }
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const double inc = coef + di * (hi - s);
            if (valid && sub == 0) e[row] = e_zero ? inc : (TM::ld(e + row) + inc);
        }
    });
    G.sync();
}

// one fused damped-Jacobi step (kernel correction through coef); returns Axi'ealt
template <class TM>
__device__ double p_jacobi(TM& G, const LevelDev& L, const PSlice& S, int tpr, const double* r, const double* ecur, double* ealt,
                           double coef, bool e_zero, bool stage = false) {
    double part = 0.0, dummy = 0.0;
    const double* xs = p_stage_x(G, e_zero ? nullptr : ecur, L.N, stage);
    with_tpr(tpr, [&](auto T) {
        constexpr int TPR = decltype(T)::value;
        constexpr int U = TM::U, RP = TM::T / TPR;
        const int sub = threadIdx.x % TPR;
        for (int base = S.first(U * RP, G.blk()); base < S.r1; base += S.step(U * RP, G.nblk())) {
            const int row0 = base + threadIdx.x / TPR;
            double d[U];
            if (xs != nullptr) row_dots<TPR, U, TM, true>(S, xs, row0, RP, sub, d);
            else if (!e_zero) row_dots<TPR, U, TM>(S, ecur, row0, RP, sub, d);
            else {
#pragma unroll
                for (int u = 0; u < U; ++u) d[u] = 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = row0 + u * RP;
                if (row < S.r1 && sub == 0) {
                    const double axi = L.Axi[row], di = L.dinv[row], ri = TM::ld(r + row), ei = e_zero ? 0.0 : TM::ld(ecur + row);
                    const double gi = ri - d[u];
                    const double en = ei + coef + di * (gi - axi * coef);
                    ealt[row] = en;
                    part = fma(axi, en, part);
                }
            }
        }
    });
    G.sum2(part, dummy);
    return part;
}

// y (+)= M x on the block's rows of M; returns sum(y) and wvec'y (wvec optional)
template <class TM>
__device__ void p_spmv(TM& G, const PSlice& S, int tpr, const double* x, double* y, bool add, const double* wvec,
                       double& sy, double& swy) {
    sy = 0.0; swy = 0.0;
    with_tpr(tpr, [&](auto T) {
        constexpr int TPR = decltype(T)::value;
        constexpr int U = TM::U, RP = TM::T / TPR;
        const int sub = threadIdx.x % TPR;
        for (int base = S.first(U * RP, G.blk()); base < S.r1; base += S.step(U * RP, G.nblk())) {
            const int row0 = base + threadIdx.x / TPR;
            double d[U];
            row_dots<TPR, U, TM>(S, x, row0, RP, sub, d);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = row0 + u * RP;
                if (row < S.r1 && sub == 0) {
                    const double v = (add ? TM::ld(y + row) : 0.0) + d[u];
                    y[row] = v; sy += v;
                    swy = fma(wvec ? wvec[row] : 0.0, v, swy);
                }
            }
        }
    });
    G.sum2(sy, swy);
}

// y (+)= B x, dense row-major n x n (n <= 2048), one warp per row, all loads of a row in flight at once.  xs: a
// 2048-double shared-memory buffer of the team (or null): x is read once per CTA instead of once per row.
template <class TM>
__device__ void p_dense(TM& G, int n, const double* __restrict__ B, const double* x, double* y, bool add) {
    const int lane = threadIdx.x & 31;
    const int gw = (G.blk() * TM::T + (int)threadIdx.x) >> 5, nw = G.nblk() * (TM::T / 32);
    double* xs = G.xs;
    if (xs != nullptr) {
        for (int j = threadIdx.x; j < n; j += TM::T) xs[j] = TM::ld(x + j);
        __syncthreads();
    }
    for (int row = gw; row < n; row += nw) {
        const double* Br = B + (size_t)row * n;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j0 = 0; j0 < n; j0 += 256) {
            double bv[8], xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 32 + lane;
                bv[u] = (j < n) ? Br[j] : 0.0;
                xv[u] = (j < n) ? (xs != nullptr ? xs[j] : TM::ld(x + j)) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u & 3] = fma(bv[u], xv[u], acc[u & 3]);
        }
        const double s = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
        if (lane == 0) y[row] = add ? (TM::ld(y + row) + s) : s;
    }
    G.sync();
}

template <class TM>
__device__ __forceinline__ void persist_body(const PersistArgs& a, TM& G, unsigned char* p_dsm, LevelDev* sl, PLevelS* ps, PState* pst) {
    constexpr int kPT = TM::T;
    const int kd = a.kd;                                   // levels 0..kd-1 explicit, level kd = dense leaf
    for (int t = threadIdx.x; t <= kd && t < kPLevels; t += kPT) sl[t] = a.levels[t];
    __syncthreads();
    // ---- stage this block's row slices of the level matrices (most visited levels first)
    {
        size_t used = 0;
        for (int t = kd; t >= 0; --t) {
            const LevelDev& L = sl[t];
            if (t < kd) stage_slice(G, &ps[t].A, L.N, L.ap, L.ai, L.av, p_dsm, used, a.smem_budget);
            else        stage_slice(G, &ps[t].A, L.N, L.ap, L.ai, L.av, p_dsm, used, 0);              // leaf: residual only, in place
            if (t < kd) stage_slice(G, &ps[t].Pu, L.N, sl[t + 1].pp, sl[t + 1].pi, sl[t + 1].pv, p_dsm, used, a.smem_budget);
            if (t >= 1) stage_slice(G, &ps[t].Td, L.N, L.tp, L.ti, L.tv, p_dsm, used, a.smem_budget);
        }
        __syncthreads();
    }
    const bool lead = (G.blk() == 0 && threadIdx.x == 0);
    // Per-level control state of the cycle (phase machine instead of recursion).  It lives in SHARED memory, not in
    // per-thread arrays: dynamically indexed thread-local arrays are local memory, and every barrier of this kernel
    // invalidates the L1 (grid.sync / barrier.cluster imply CCTL.IVALL), so each access would be an L2 round trip.
    // Every thread executes the same control flow and writes the same values; an iteration first snapshots what it
    // needs into registers, then passes a block barrier, and writes only after it -- so no thread can overwrite a
    // value another thread has not read yet, and a late write only repeats the value that is already there.
    PState& st = *pst;
    for (int t = threadIdx.x; t < kPLevels; t += kPT) { st.phase[t] = 0; st.zero[t] = 1; st.ecur[t] = nullptr; st.ealt[t] = nullptr; st.sum_r[t] = 0.0; st.dot_e[t] = 0.0; }
    __syncthreads();
    const long long t_kernel0 = clock64();
    const int tpr0 = kd > 0 ? a.tpr[0] : 32;
    int dbg_level__ = 0; (void)dbg_level__;

    // r = b - A*x ; res0 = norm(r)                                      Class_AMG.m:89
    double s1, s2;
    PDBG(25, p_resid(G, ps[0].A, tpr0, a.b, a.x, sl[0].r, true, s1, s2));
    const double res0 = sqrt(s2);
    double sum_r0 = s1;
    double res_prev = res0, rel_prev = 1.0, rel_res = 0.0;
    int it = 0, hist = 1;
    if (lead) { a.relk[0] = 1.0; a.rho[0] = NAN; }
    if (res0 == 0.0) {
        if (lead) { a.relk[0] = 0.0; a.rho[0] = INFINITY; a.it_out[0] = 0; a.it_out[1] = 1; a.it_out[2] = 0; }
        return;
    }
    it = 1;
    while (rel_prev > a.retol && it <= a.maxit) {                       // Class_AMG.m:95
        // ---------------- one cycle on level 0: rhs sl[0].r -> ecur[0]
        __syncthreads();
        for (int t = threadIdx.x; t <= kd; t += kPT) { st.ecur[t] = sl[t].e; st.ealt[t] = sl[t].pcg; }
        if (threadIdx.x == 0) { st.phase[0] = 0; st.zero[0] = 1; st.sum_r[0] = sum_r0; }
        __syncthreads();
        int k = 0;
        double* e_top = nullptr;
        while (true) {
            const LevelDev& L = sl[k];
            const PLevelS& P = ps[k];
            dbg_level__ = k;
            // snapshot of this level's state (and of the child's correction), then the barrier that orders it before any write
            const int ph = st.phase[k];
            const bool zk = st.zero[k] != 0;
            double* ec = st.ecur[k]; double* ea = st.ealt[k];
            const double sr = st.sum_r[k], de = st.dot_e[k];
            double* ec_child = (k < kd) ? st.ecur[k + 1] : nullptr;
            __syncthreads();
            if (k == kd) {                                              // dense tail operator
                if (zk) PDBG(29, p_dense(G, L.N, L.B, L.r, L.e, false));
                else { double d1, d2; PDBG(25, p_resid(G, P.A, 32, L.r, L.e, L.g, false, d1, d2)); PDBG(29, p_dense(G, L.N, L.B, L.g, L.e, true)); }
                st.ecur[k] = L.e;
                if (k == 0) { e_top = L.e; break; }
                --k; continue;
            }
            const int tpr = a.tpr[k];
            if (ph == 0 || ph == 3) {                                   // pre- (0) or post-smoothing (3)
                const int post = (ph == 3) ? 1 : 0;
                bool ez = (ph == 0) ? zk : false;
                if (a.smoth == 0 && ez) {
                    for (int i = P.A.first(kPT, G.blk()) + threadIdx.x; i < P.A.r1; i += P.A.step(kPT, G.nblk())) ec[i] = 0.0;
                    G.sync();
                }
                double dotAe = ez ? 0.0 : de;
                if (L.bigph) {
                    for (int s = 0; s < a.smoth; ++s) {
                        double sg, sg2;
                        PDBG(25, p_resid(G, P.A, tpr, L.r, ez ? nullptr : ec, L.g, a.isnsp != 0, sg, sg2, a.xs_ok[k] ? L.N : 0));
                        const double coef = a.isnsp ? sg / L.xx : 0.0;
                        PDBG(26, p_gs_apply(G, L, P.A, a.tpr_gs > 0 ? a.tpr_gs : tpr, L.g, ec, coef, post, ez));
                        ez = false;
                    }
                } else {
                    for (int s = 0; s < a.smoth; ++s) {
                        const double coef = a.isnsp ? (sr - dotAe) / L.xx : 0.0;
                        PDBG(27, dotAe = p_jacobi(G, L, P.A, tpr, L.r, ec, ea, coef, ez, a.xs_ok[k] != 0));
                        double* t = ec; ec = ea; ea = t;
                        ez = false;
                    }
                }
                if (post) {                                             // this level's visit is complete
                    st.ecur[k] = ec; st.ealt[k] = ea; st.dot_e[k] = dotAe;
                    if (k == 0) { e_top = ec; break; }
                    --k; continue;
                }
                // restriction: r_{k+1} = Pro' (r - A e)                  MG_Wcycle.m:26
                double d1, d2;
                PDBG(25, p_resid(G, P.A, tpr, L.r, (ez ? nullptr : ec), L.g, false, d1, d2, a.xs_ok[k] ? L.N : 0));
                PDBG(28, p_spmv(G, ps[k + 1].Td, a.tpr_p[k + 1], L.g, sl[k + 1].r, false, nullptr, d1, d2));
                st.ecur[k] = ec; st.ealt[k] = ea; st.dot_e[k] = dotAe;
                st.sum_r[k + 1] = d1;
                st.phase[k] = 1; st.phase[k + 1] = 0; st.zero[k + 1] = 1; ++k; continue;
            }
            if (ph == 1 && a.wcycle && (k + 1 != a.J - 1)) {            // second coarse visit   :30
                st.phase[k] = 2; st.phase[k + 1] = 0; st.zero[k + 1] = 0; ++k; continue;
            }
            {   // prolongation e += Pro e_{k+1}, with Axi'e for the post-smoother        :32
                double d1, d2;
                PDBG(28, p_spmv(G, P.Pu, a.tpr_p[k + 1], ec_child, ec, true, L.Axi, d1, d2));
                st.dot_e[k] = d2;
                st.phase[k] = 3; continue;
            }
        }
        dbg_level__ = 0;
        // ---------------- x += e ; r = b - A*x ; res = norm(r)          Class_AMG.m:96-104
        for (int i = ps[0].A.first(kPT, G.blk()) + threadIdx.x; i < ps[0].A.r1; i += ps[0].A.step(kPT, G.nblk())) a.x[i] = __ldcg(a.x + i) + __ldcg(e_top + i);
        G.sync();
        PDBG(25, p_resid(G, ps[0].A, tpr0, a.b, a.x, sl[0].r, true, s1, s2));
        sum_r0 = s1;
        const double res = sqrt(s2);
        rel_res = res / res0;
        const double rho = res / res_prev;
        if (lead) { a.relk[it] = rel_res; a.rho[it] = rho; }
        res_prev = res; rel_prev = rel_res;
        ++it; ++hist;
        if (rho > 1.0) break;                                           // Class_AMG.m:106
    }
    if (lead) { a.it_out[0] = it - 1; a.it_out[1] = hist; a.it_out[2] = 0; g_pdbg[7 * 16] += (unsigned long long)(clock64() - t_kernel0); g_pdbg[128 + 7 * 16] += 1ull; }
    (void)rel_res;
}

__global__ void __launch_bounds__(kPT, kPBlocksPerSM) persist_solve_kernel(const PersistArgs a) {
    extern __shared__ __align__(16) unsigned char p_dsm[];
    __shared__ double red[128];
    __shared__ LevelDev sl[kPLevels];
    __shared__ PLevelS ps[kPLevels];
    __shared__ PState pst;
    // the first kGridXs doubles of the dynamic shared memory hold the copy of a gathered vector, the rest the staged slices
    GridTeam G{cg::this_grid(), a.part, red, 0, reinterpret_cast<double*>(p_dsm), kGridXs};
    persist_body(a, G, p_dsm + sizeof(double) * kGridXs, sl, ps, &pst);
}

// The same solve loop inside ONE thread-block cluster (launched with a runtime cluster dimension of 16, or 8 where
// the non-portable size is not available): with the whole late-phase hierarchy (N ~ 3e4, 1e5 nonzeros per level)
// the dependent passes are latency, not throughput, so 16 SMs are enough and the hardware cluster barrier
// (~0.2 us) replaces the grid-wide barrier (>= 1.2 us + the skew of 296 blocks) and the L2 round trip of the
// reduction partials; each CTA keeps its row slices of the level matrices in its 200 KB of shared memory.
__global__ void __launch_bounds__(kCTT, 1) cluster_solve_kernel(const PersistArgs a) {
    extern __shared__ __align__(16) unsigned char p_dsm[];
    __shared__ double red[128];
    __shared__ double slots[2 * 2 * kCMax];
    __shared__ LevelDev sl[kPLevels];
    __shared__ PLevelS ps[kPLevels];
    cg::cluster_group cl = cg::this_cluster();
    __shared__ double xs[2048];
    __shared__ PState pst;
    ClusterTeam G{red, slots, (int)cl.block_rank(), (int)cl.num_blocks(), 0, xs};
    persist_body(a, G, p_dsm, sl, ps, &pst);
    cluster_barrier();                                     // no CTA exits while a peer may still write into its shared memory
}

template <class F>
void dispatch_tpr(double avg, F&& f) {
    if (avg <= 3.0) f(std::integral_constant<int, 2>());
    else if (avg <= 6.0) f(std::integral_constant<int, 4>());
    else if (avg <= 12.0) f(std::integral_constant<int, 8>());
    else if (avg <= 24.0) f(std::integral_constant<int, 16>());
    else f(std::integral_constant<int, 32>());
}

int grid_rows(int n, int tpr) {
    int g = cdiv((int64_t)n * tpr, 256);
    if (g > kMaxPartBlocks) g = kMaxPartBlocks;
    if (g < 1) g = 1;
    return g;
}

// g = r - A e with partials; returns the number of partial pairs
int launch_resid(ssn_ctx* c, const Level& L, const double* r, const double* e, double* g, double* part) {
    const double avg = L.N ? (double)L.A.nnz / L.N : 0.0;
    int np = 1;
    dispatch_tpr(avg, [&](auto T) {
        constexpr int TPR = decltype(T)::value;
        np = grid_rows(L.N, TPR);
        SSN_LAUNCH(c, resid_kernel<TPR>, np, 256, 0, L.N, L.A.ptr.p, L.A.idx.p, L.A.val.p, r, e, g, part);
    });
    return np;
}

void launch_apply(ssn_ctx* c, const LevelDev& Ld, const Level& L, const double* g, double* e, const double* part, int np,
                  int isnsp, int post, bool e_zero) {
    const double avg = L.N ? (double)L.A.nnz / L.N : 0.0;
    dispatch_tpr(avg, [&](auto T) {
        constexpr int TPR = decltype(T)::value;
        const int gr = grid_rows(L.N, TPR);
        SSN_LAUNCH(c, smooth_apply_kernel<TPR>, gr, 256, 0, Ld, g, e, part, np, isnsp, post, e_zero ? 1 : 0);
    });
}

LevelDev level_dev(const Level& L) {
    LevelDev d{};
    d.N = L.N; d.Nf = L.Nf; d.bigph = L.bigph;
    d.ap = L.A.ptr.p; d.ai = L.A.idx.p; d.av = L.A.val.p;
    d.pp = L.P.ptr.p; d.pi = L.P.idx.p; d.pv = L.P.val.p;
    d.tp = L.Pt.ptr.p; d.ti = L.Pt.idx.p; d.tv = L.Pt.val.p;
    d.dinv = L.dinv.p; d.Axi = L.Axi.p; d.xx = L.xx;
    d.r = L.r.p; d.e = L.e.p; d.g = L.g.p; d.pcg = L.pcg.p; d.B = L.B.p;
    return d;
}

constexpr int kPartR = 2 * kMaxPartBlocks;                 // H.part layout: [resid pairs | sum(r) | dot A | dot B]
constexpr int kPartDotA = 3 * kMaxPartBlocks;
constexpr int kPartDotB = 4 * kMaxPartBlocks;

void smooth_host(ssn_ctx* c, Hierarchy& H, int k, int isnsp, int post, bool e_zero) {
    Level& L = H.lv[k];
    if (H.smoth == 0) { if (e_zero) fill_double(c, L.e, L.N, 0.0); return; }
    if (L.bigph) {                                        // block Gauss-Seidel level: residual + coupled update
        const LevelDev Ld = level_dev(L);
        for (int it = 0; it < H.smoth; ++it) {
            const int np = launch_resid(c, L, L.r, e_zero ? nullptr : L.e.p, L.g, H.part);
            launch_apply(c, Ld, L, L.g, L.e, H.part, np, isnsp, post, e_zero);
            e_zero = false;
        }
        return;
    }
    const double avg = L.N ? (double)L.A.nnz / L.N : 0.0;
    double* part = H.part.p;
    const int nb1 = std::min(kMaxPartBlocks, std::max(1, cdiv(L.N, 256)));
    int np_d = 0;
    if (isnsp) {
        if (!post) SSN_LAUNCH(c, dot_parts_kernel, nb1, 256, 0, L.N, L.r.p, nullptr, part + kPartR);   // sum(r): once per visit
        if (!e_zero) { SSN_LAUNCH(c, dot_parts_kernel, nb1, 256, 0, L.N, L.Axi.p, L.e.p, part + kPartDotA); np_d = nb1; }
    }
    double* dot_in = part + kPartDotA; double* dot_out = part + kPartDotB;
    for (int it = 0; it < H.smoth; ++it) {
        const LevelDev Ld = level_dev(L);
        dispatch_tpr(avg, [&](auto T) {
            constexpr int TPR = decltype(T)::value;
            const int gr = grid_rows(L.N, TPR);
            SSN_LAUNCH(c, smooth_fused_kernel<TPR>, gr, 256, 0, Ld, L.r.p, L.e.p, L.pcg.p, part + kPartR, nb1, dot_in, np_d, dot_out,
                       isnsp, e_zero ? 1 : 0);
            np_d = gr;
        });
        std::swap(L.e.p, L.pcg.p);                        // ping-pong: the smoothed iterate is the level's e again
        std::swap(dot_in, dot_out);
        e_zero = false;
    }
}

// cycle on level k (0-based): rhs in lv[k].r, correction in lv[k].e
ClusterPlan plan_for_level(const Hierarchy& H, int k) {
    ClusterPlan q = H.cluster_plan;
    const int sh = k - H.cluster_from;
    if (sh > 0) {
        q.nl -= sh;
        for (int t = 0; t + sh < kCL; ++t) { q.staged[t] = q.staged[t + sh]; q.smem_off[t] = q.smem_off[t + sh]; q.rpc[t] = q.rpc[t + sh]; q.cap[t] = q.cap[t + sh]; }
    }
    return q;
}

// Builds B_k for the tail levels (coarsest first); (isnsp, cycle) are part of the operator.
void build_dense_tail(ssn_ctx* c, Hierarchy& H, int isnsp, bool wcycle) {
    H.dense_isnsp = isnsp; H.dense_w = wcycle ? 1 : 0; H.dense_from = H.J;
    if (!c->dense_tail) return;
    const int J = H.J;
    int from = J;
    for (int k = J - 1; k >= 0; --k) {
        if (H.lv[k].N <= c->dense_max_n && !H.lv[k].bigph && H.lv[k].N > 0) from = k; else break;
    }
    if (from >= J) return;
    Phase ph(c, "solve.build_dense_tail");
    for (int k = J - 1; k >= from; --k) {
        Level& L = H.lv[k];
        const int N = L.N;
        L.B.alloc(c, (size_t)N * N);
        const LevelDev Ld = level_dev(L);
        if (k == J - 1) {
            if (N <= 32) SSN_LAUNCH(c, dense_coarsest_warp_kernel, N, 32, 0, Ld, L.B.p);
            else {
                const size_t smem = sizeof(double) * 4 * (size_t)N;
                SSN_LAUNCH(c, dense_coarsest_block_kernel, N, 256, smem, Ld, L.B.p);
            }
            continue;
        }
        Level& Lc = H.lv[k + 1];
        const LevelDev Lcd = level_dev(Lc);
        const int twice = (wcycle && (k + 1 != J - 1)) ? 1 : 0;
        // columns per CTA: the fewest that still cover the level in ONE wave of CTAs (a CTA's time grows with its columns,
        // and a level of 651 rows at 8 columns per CTA would use 82 of the 148 SMs)
        auto stride_of = [](int cc) { return cc == 1 ? 1 : (cc % 2 == 0 ? cc + 1 : cc + 2); };
        auto smem_for = [&](int cc) { return sizeof(double) * (size_t)stride_of(cc) * (3 * (size_t)N + 3 * (size_t)Lc.N); };
        int C = 8;
        for (int cc : {1, 2, 3, 4, 5, 6, 8}) if (cdiv(N, cc) <= c->num_sms) { C = cc; break; }
        while (C > 1 && smem_for(C) > 200 * 1024) C = (C == 8) ? 6 : C - 1;
        const size_t smem = smem_for(C);
        const int grid = cdiv(N, C);
        auto go = [&](auto T) {
            constexpr int CC = decltype(T)::value;
            SSN_CUDA(cudaFuncSetAttribute(dense_build_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SSN_LAUNCH(c, dense_build_kernel<CC>, grid, kDT, smem, Ld, Lcd, Lc.B.p, H.smoth, isnsp, twice, L.B.p);
        };
        switch (C) {
            case 8: go(std::integral_constant<int, 8>()); break;
            case 6: go(std::integral_constant<int, 6>()); break;
            case 5: go(std::integral_constant<int, 5>()); break;
            case 4: go(std::integral_constant<int, 4>()); break;
            case 3: go(std::integral_constant<int, 3>()); break;
            case 2: go(std::integral_constant<int, 2>()); break;
            default: go(std::integral_constant<int, 1>()); break;
        }
    }
    H.dense_from = from;
    H.hdev.resize(J);
    for (int k = 0; k < J; ++k) H.hdev[k] = level_dev(H.lv[k]);
    upload_small(c, H.dev.p, H.hdev.data(), sizeof(LevelDev) * J);
}

void cycle_host(ssn_ctx* c, Hierarchy& H, int k, int isnsp, bool wcycle, bool e_zero) {
    if (H.dense_isnsp != isnsp || H.dense_w != (wcycle ? 1 : 0)) build_dense_tail(c, H, isnsp, wcycle);
    if (k >= H.dense_from) {
        Level& L = H.lv[k];
        Phase ph(c, "solve.dense_apply");
        const int grid = cdiv((int64_t)L.N * 32, 256);
        if (e_zero) SSN_LAUNCH(c, dense_apply_kernel, grid, 256, 0, L.N, L.B.p, L.r.p, L.e.p, 0);
        else {
            launch_resid(c, L, L.r, L.e.p, L.g, H.part);
            SSN_LAUNCH(c, dense_apply_kernel, grid, 256, 0, L.N, L.B.p, L.g.p, L.e.p, 1);
        }
        return;
    }
    if (k >= H.cluster_from && k < H.J - 1) {
        Phase ph(c, "solve.cluster_cycle_kernel");
        SSN_LAUNCH(c, cluster_cycle_kernel, kCS, kCT, H.cluster_smem, H.dev.p, k, H.J, H.smoth, isnsp, wcycle ? 1 : 0,
                   e_zero ? 1 : 0, plan_for_level(H, k));
        return;
    }
    if (k >= H.small_from || k == H.J - 1) {
        SSN_REQUIRE(H.J - k <= 16, SSN_E_INVALID, "hierarchy deeper than 16 small levels");
        Phase ph(c, "solve.coarse_cycle_kernel");
        SSN_LAUNCH(c, coarse_cycle_kernel, 1, kCycleThreads, 0, H.dev.p, k, H.J, H.smoth, isnsp, wcycle ? 1 : 0, e_zero ? 1 : 0);
        return;
    }
    Level& L = H.lv[k]; Level& Lc = H.lv[k + 1];
    smooth_host(c, H, k, isnsp, 0, e_zero);
    launch_resid(c, L, L.r, (e_zero && H.smoth == 0) ? nullptr : L.e.p, L.g, H.part);
    spmv(c, Lc.Pt, L.g, Lc.r);                                            // MG_Wcycle.m:26
    cycle_host(c, H, k + 1, isnsp, wcycle, true);                         // :28
    if (wcycle) cycle_host(c, H, k + 1, isnsp, wcycle, false);            // :30
    spmv_add(c, Lc.P, Lc.e, L.e);                                         // :32
    smooth_host(c, H, k, isnsp, 1, false);                                // :34-42
}

int tpr_for(double avg) { return avg <= 3.0 ? 2 : (avg <= 6.0 ? 4 : (avg <= 12.0 ? 8 : (avg <= 24.0 ? 16 : 32))); }

// Runs Class_AMG's solve loop in ONE kernel: inside a single thread-block cluster when the hierarchy is small
// enough for 16 SMs (late-phase systems: always), else as the grid-wide cooperative kernel.  Returns false when the
// hierarchy does not qualify (no dense tail / too many large levels), in which case the caller launches kernel by kernel.
bool persist_solve(ssn_ctx* c, Hierarchy& H, const double* b, double* x, const AmgOptions& o, bool wcycle, int& it,
                   double& rel_res, std::vector<double>& relk, std::vector<double>& rho) {
    if (H.dense_from >= H.J || H.dense_from >= kPLevels) return false;
    int64_t nnz = 0;
    for (int k = 0; k < H.dense_from; ++k) nnz += H.lv[k].A.nnz;
    // the persistent kernels trade occupancy for latency: with large level matrices (early SsN steps,
    // millions of nonzeros) the cycle is bandwidth-bound and the multi-block kernels win
    if (nnz > c->persist_max_nnz) return false;
    for (int k = 0; k < H.dense_from; ++k) if (H.lv[k].bigph && k != 0) return false;
    PersistArgs a{};
    a.levels = H.dev.p; a.J = H.J; a.kd = H.dense_from; a.smoth = H.smoth; a.isnsp = o.isnsp; a.wcycle = wcycle ? 1 : 0;
    a.b = b; a.x = x; a.retol = o.retol; a.maxit = o.maxit;
    for (int k = 0; k <= H.dense_from && k < kPLevels; ++k) {
        const Level& L = H.lv[k];
        a.tpr[k] = tpr_for(L.N ? (double)L.A.nnz / L.N : 0.0);
        a.tpr_p[k] = (k > 0) ? tpr_for(L.N ? (double)L.P.nnz / std::max(1, L.N) : 0.0) : 2;
    }
    for (int k = 0; k < kPLevels; ++k) a.xs_ok[k] = 0;
    for (int k = 0; k < H.dense_from && k < kPLevels; ++k) {
        const Level& L = H.lv[k];
        a.xs_ok[k] = (c->stage_dense && L.N > 0 && L.N <= kGridXs && (double)L.A.nnz / L.N >= 48.0) ? 1 : 0;
    }
    const int hl = o.maxit + 2;
    Buf<double> hist(c, (size_t)2 * hl);
    Buf<int> iout(c, 4);
    a.relk = hist.p; a.rho = hist.p + hl; a.it_out = iout.p;
    bool launched = false, have_hi = false;
    int hi[4];
    // ---- one cluster, level vectors in distributed shared memory (amg_cluster.cu); it reports a level-1 matrix whose
    // diagonal blocks are not diagonal (the two-half-sweep smoother does not apply) without having changed x
    if (c->cluster_solve && c->dsm_solve && nnz <= c->cluster_max_nnz && dsm_cluster_solve(c, H, b, x, o, wcycle, hist.p, hl, iout.p)) {
        read_back(c, iout.p, hi, 4);
        if (hi[2] == 0) {
            launched = true; have_hi = true; c->last_dsm_halo = hi[3];
            if (c->prof) c->prof_acc["dsm_solve_kernel: halo entries of CTA 0 over all levels = " + std::to_string(hi[3])].second += 1;
        }
    }
    if (!launched && c->cluster_solve && nnz <= c->cluster_max_nnz) {
        // ---- one cluster: 16 CTAs (non-portable size) where the device can co-schedule them, else 8
        static int cluster_ctas = -1;                       // probed once per process
        const size_t smem = (size_t)200 * 1024;              // + 22 KB static (reduction slots, level tables, the dense tail input)
        if (cluster_ctas < 0) {
            cluster_ctas = 0;
            if (cudaFuncSetAttribute(cluster_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
                const bool np_ok = cudaFuncSetAttribute(cluster_solve_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
                for (int want : {np_ok ? kCMax : 8, 8}) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(want); cfg.blockDim = dim3(kCTT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = want; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    int nclusters = 0;
                    if (cudaOccupancyMaxActiveClusters(&nclusters, cluster_solve_kernel, &cfg) == cudaSuccess && nclusters >= 1) { cluster_ctas = want; break; }
                }
            }
            (void)cudaGetLastError();
            { const char* e = getenv("SSN_CLUSTER_CTAS"); if (e && (atoi(e) == 8 || atoi(e) == 4 || atoi(e) == 2) && cluster_ctas >= atoi(e)) cluster_ctas = atoi(e); }
        }
        if (cluster_ctas > 0) {
            Phase ph(c, "solve.cluster_solve_kernel");
            a.part = nullptr;
            a.smem_budget = smem;
            { const char* e = getenv("SSN_PERSIST_SMEM"); if (e && e[0] == '0') a.smem_budget = 0; }
            // lanes per row: as few as keep every row of a CTA's slice in flight at once -- with 16 SMs a pass is
            // bound by the dependent L2 round trips of its row loop, so one pass over the slice beats wide rows
            auto fit = [&](int t, int rows, int u) { const int per = (rows + cluster_ctas - 1) / cluster_ctas; while (t > 1 && (int64_t)per * t > (int64_t)kCTT * u) t >>= 1; return t; };
            for (int k = 0; k <= H.dense_from && k < kPLevels; ++k) {
                a.tpr[k] = fit(a.tpr[k], H.lv[k].N, ClusterTeam::U);
                if (k > 0) a.tpr_p[k] = fit(a.tpr_p[k], H.lv[k - 1].N, ClusterTeam::U);
            }
            a.tpr_gs = fit(a.tpr[0], H.lv[0].N, 1);         // the coupled block Gauss-Seidel update walks one row per thread group
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cluster_ctas); cfg.blockDim = dim3(kCTT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            SSN_CUDA(cudaLaunchKernelEx(&cfg, cluster_solve_kernel, a));
            c->launches++;
            launched = true;
        }
    }
    Buf<double> part;
    if (!launched) {
        Phase ph(c, "solve.persist_solve_kernel");
        const size_t smem = (size_t)(200 / kPBlocksPerSM - 6) * 1024;
        SSN_CUDA(cudaFuncSetAttribute(persist_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        SSN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persist_solve_kernel, kPT, smem));
        if (per_sm < kPBlocksPerSM) return false;
        const int grid = c->num_sms * kPBlocksPerSM;
        a.smem_budget = smem - sizeof(double) * kGridXs;      // what is left for staged matrix slices
        { const char* e = getenv("SSN_PERSIST_SMEM"); if (e && e[0] == '0') a.smem_budget = 0; }
        part.alloc(c, (size_t)4 * grid);
        a.part = part.p;
        void* args[] = {&a};
        SSN_CUDA(cudaLaunchCooperativeKernel((void*)persist_solve_kernel, dim3(grid), dim3(kPT), args, smem, c->stream));
        c->launches++;
    }
    if (!have_hi) read_back(c, iout.p, hi, 4);
    it = hi[0];
    const int len = hi[1];
    std::vector<double> hh((size_t)2 * hl);
    read_back(c, hist.p, hh.data(), hh.size());
    relk.assign(hh.begin(), hh.begin() + len);
    rho.assign(hh.begin() + hl, hh.begin() + hl + len);
    rel_res = (len > 1) ? relk[len - 1] : 0.0;            // no cycle ran (zero right-hand side): Class_AMG's rel_res stays 0
    return true;
}

}  // namespace

// Builds the shared-memory staging plan of the cluster kernel for levels k0..J-1 (once per hierarchy).
void build_cluster_plan(ssn_ctx* c, Hierarchy& H) {
    H.cluster_from = H.J;                                  // disabled unless everything below qualifies
    const int J = H.J;
    if (J < 2 || H.lv[J - 1].N > 32) return;
    int k0 = J - 1;
    while (k0 > 0 && H.lv[k0 - 1].N <= 4096 && H.lv[k0 - 1].A.nnz <= (1 << 18) && !H.lv[k0 - 1].bigph && (J - (k0 - 1)) <= kCL) --k0;
    if (k0 >= J - 1) return;
    Buf<int> bounds(c, (size_t)J * (kCS + 1));
    SSN_LAUNCH(c, slice_bounds_kernel, cdiv(J * (kCS + 1), 128), 128, 0, H.dev.p, J, bounds.p);
    std::vector<int> hb((size_t)J * (kCS + 1));
    read_back(c, bounds.p, hb.data(), hb.size());
    ClusterPlan plan{};
    plan.nl = J - k0;
    size_t budget = 200 * 1024, used = 0;
    for (int t = 0; t < kCL; ++t) { plan.staged[t] = 0; plan.smem_off[t] = 0; plan.rpc[t] = 1; plan.cap[t] = 0; }
    for (int k = J - 1; k >= k0; --k) {                    // coarsest levels first: they are visited most
        const int t = k - k0;
        const int N = H.lv[k].N, rpc = (N + kCS - 1) / kCS;
        plan.rpc[t] = rpc;
        int cap = 0;
        for (int r = 0; r < kCS; ++r) cap = std::max(cap, hb[(size_t)k * (kCS + 1) + r + 1] - hb[(size_t)k * (kCS + 1) + r]);
        cap = ((cap + 1) / 2) * 2;
        const size_t bytes = (size_t)(((rpc + 1) * 4 + 15) / 16) * 16 + (size_t)cap * 12 + 16;
        if (used + bytes <= budget) { plan.staged[t] = 1; plan.smem_off[t] = (int)used; plan.cap[t] = cap; used += ((bytes + 15) / 16) * 16; }
    }
    H.cluster_plan = plan; H.cluster_smem = used; H.cluster_from = k0;
    SSN_CUDA(cudaFuncSetAttribute(cluster_cycle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(used ? used : 16)));
}

// one cluster of 16 (or 8) CTAs x 1024 threads: which = 2 release/acquire barrier alone, 3 a global store per thread
// before every barrier, 4 relaxed arrive (no MEMBAR) + wait, 5 store + barrier + a dependent L2 gather (ld.cg) per thread,
// 6 as 5 with a plain (L1-cached) gather
__global__ void __launch_bounds__(1024, 1) cluster_barrier_bench_kernel(double* buf, int iters, int which, long long* cycles_out) {
    cg::cluster_group cl = cg::this_cluster();
    const int gt = (int)cl.block_rank() * 1024 + threadIdx.x, nt = (int)cl.num_blocks() * 1024;
    cluster_barrier();
    const long long t0 = clock64();
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        if (which == 3 || which >= 5) buf[gt] = acc + i;
        if (which == 4) asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
        else cluster_barrier();
        if (which == 5) acc += __ldcg(buf + ((gt * 7 + i * 131) % nt));
        if (which == 6) acc += buf[(gt * 7 + i * 131) % nt];
    }
    if (cl.block_rank() == 0 && threadIdx.x == 0) cycles_out[0] = clock64() - t0;
    if (acc == 123.456) buf[0] = acc;
    cluster_barrier();
}

// cycles per grid barrier: which = 0 cooperative-groups grid.sync(), 1 = grid_barrier() (development aid)
double barrier_bench(ssn_ctx* c, int iters, int which) {
    if (which >= 2) {
        Buf<double> buf(c, 16 * 1024); buf.zero();
        Buf<long long> out(c, 1);
        (void)cudaFuncSetAttribute(cluster_barrier_bench_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int want : {16, 8}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(want); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = want; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (cudaLaunchKernelEx(&cfg, cluster_barrier_bench_kernel, buf.p, iters, which, out.p) == cudaSuccess) {
                long long cyc = read_scalar(c, out.p);
                return (double)cyc / (double)iters;
            }
            (void)cudaGetLastError();
        }
        return -1.0;
    }
    Buf<unsigned> bar(c, 64); bar.zero();
    Buf<long long> out(c, 1);
    unsigned* bp = bar.p; long long* op = out.p;
    void* args[] = {&bp, &iters, &which, &op};
    SSN_CUDA(cudaLaunchCooperativeKernel((void*)barrier_bench_kernel, dim3(c->num_sms * kPBlocksPerSM), dim3(kPT), args, 0, c->stream));
    long long cyc = read_scalar(c, out.p);
    return (double)cyc / (double)iters;
}

void debug_cycles_persist(unsigned long long* out256, bool reset) {
    cudaMemcpyFromSymbol(out256, g_pdbg, sizeof(unsigned long long) * 256);
    if (reset) { unsigned long long z[256] = {0}; cudaMemcpyToSymbol(g_pdbg, z, sizeof(z)); }
}

void debug_cycles(unsigned long long* out64, bool reset) {
    cudaMemcpyFromSymbol(out64, g_dbg_cycles, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_dbg_cycles, z, sizeof(z)); }
}

void mg_cycle(ssn_ctx* c, const double* r_dev, int isnsp, int k1, double* e_dev, bool wcycle, bool e_is_zero) {
    SSN_REQUIRE(c->hier != nullptr, SSN_E_NO_HIERARCHY, "MG cycle called without a live hierarchy");
    Hierarchy& H = *c->hier;
    SSN_REQUIRE(k1 >= 1 && k1 <= H.J, SSN_E_INVALID, "MG cycle: level out of range");
    Level& L = H.lv[k1 - 1];
    SSN_CUDA(cudaMemcpyAsync(L.r.p, r_dev, sizeof(double) * L.N, cudaMemcpyDeviceToDevice, c->stream));
    if (!e_is_zero) SSN_CUDA(cudaMemcpyAsync(L.e.p, e_dev, sizeof(double) * L.N, cudaMemcpyDeviceToDevice, c->stream));
    cycle_host(c, H, k1 - 1, isnsp, wcycle, e_is_zero);
    SSN_CUDA(cudaMemcpyAsync(e_dev, L.e.p, sizeof(double) * L.N, cudaMemcpyDeviceToDevice, c->stream));
}

// Class_AMG.m:41-110
void class_amg(ssn_ctx* c, const CsrView& A, const double* b, const AmgOptions& o, bool keep, double* x, int* it_out,
               double* rel_res_out, double* rel_resk, double* rhok, int* hist_len) {
    amg_setup(c, A, o);
    Phase ph_solve(c, "class_amg solve loop total");
    Hierarchy& H = *c->hier;
    Level& L = H.lv[0];
    const int n = L.N;
    if (c->prof) { std::string s = "levels:"; for (auto& l : H.lv) s += " " + std::to_string(l.N) + "/" + std::to_string(l.A.nnz); s += " small_from=" + std::to_string(H.small_from); c->prof_acc[s].second += 1; }
    if (o.guess) SSN_CUDA(cudaMemcpyAsync(x, o.guess, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    else fill_double(c, x, n, 0.0);
    int it = 0;
    double rel_res = 0.0;
    std::vector<double> relk(1, 1.0), rho(1, NAN);
    const bool isv = (o.cycle == 'v'), isw = (o.cycle == 'w');
    bool done = false;
    if (c->persist && (isv || isw)) {
        if (H.dense_isnsp != o.isnsp || H.dense_w != (isw ? 1 : 0)) build_dense_tail(c, H, o.isnsp, isw);
        done = persist_solve(c, H, b, x, o, isw, it, rel_res, relk, rho);
    }
    if (!done) {
    // r = b - A*x ; res0 = norm(A*x - b)                                 Class_AMG.m:89
    double h[2];
    int np = launch_resid(c, L, b, x, L.r, H.part);
    SSN_LAUNCH(c, reduce_parts_kernel, 1, 256, 0, H.part.p, np, H.scal.p);
    read_back(c, H.scal.p, h, 2);
    const double res0 = std::sqrt(h[1]);
    double res_prev = res0;
    if (res0 == 0.0) {
        rel_res = 0.0; relk.assign(1, 0.0); rho.assign(1, INFINITY);
    } else {
        it = 1;
        while (relk[it - 1] > o.retol && it <= o.maxit) {                 // Class_AMG.m:95
            if (isv || isw) {
                cycle_host(c, H, 0, o.isnsp, isw, true);                  // e = cycle(r)
                SSN_LAUNCH(c, axpy_kernel, cdiv(n, 256), 256, 0, n, 1.0, L.e.p, x);
            }
            np = launch_resid(c, L, b, x, L.r, H.part);                   // next r, and res = norm(A*x-b)
            SSN_LAUNCH(c, reduce_parts_kernel, 1, 256, 0, H.part.p, np, H.scal.p);
            read_back(c, H.scal.p, h, 2);
            const double res = std::sqrt(h[1]);
            rel_res = res / res0;
            relk.push_back(rel_res);
            rho.push_back(res / res_prev);                                // res/norm(r), r = previous residual
            res_prev = res;
            ++it;
            if (rho[it - 1] > 1.0) break;                                 // Class_AMG.m:106
        }
        relk.resize(it); rho.resize(it); --it;
    }
    }
    if (it_out) *it_out = it;
    if (rel_res_out) *rel_res_out = rel_res;
    if (hist_len) *hist_len = (int)relk.size();
    if (rel_resk) std::memcpy(rel_resk, relk.data(), sizeof(double) * relk.size());
    if (rhok) std::memcpy(rhok, rho.data(), sizeof(double) * rho.size());
    if (!keep) amg_clear(c);                                              // Class_AMG.m:110
}

// [x,it,rel_res,rel_resk,rhok] = twogrid_bigph(A,b,amg_options) -- AMG/twogrid_bigph.m:24-116: the two-level
// method behind Hybrid_twogrid (inner_solver = 5).  Setup (:26-47) is the first coarsening step of
// Class_AMG with bigph = 1 -- the same block Gauss-Seidel smoother R (:35 == Class_AMG.m:56-59), the same
// interpolation W = -Aff\Afc, row-normalised when isnsp (:44-46 == transfer.m:20-25), Ac = Pro'*A*Pro --
// so the hierarchy code builds it with max_levels = 2.  One iteration (twogrid_it, :79-116): `smoth`
// pre-smoothing steps with R (kernel correction when isnsp), restriction, coarse correction by
// PCG(Ac, rrc, {retol [] -> 1e-11, maxit 100, precd 2}) (:98-99), prolongation, `smoth` post-smoothing
// steps with R'.  The outer loop (:62-76) is Class_AMG's.
//
// generic = true is AMG/twogrid.m:1-150, the same method for a general graph Laplacian: with bigph = 0 the
// smoother is damped Jacobi 0.5*D^-1 (:59) and the coarse level comes from mis_set(A,1/4) + the standard
// interpolation W1 + 0.5*W2 (:73-92) -- the MIS step of transfer.m with theta = 1/4; with bigph = 1 it is
// twogrid_bigph after the checks of :36-38,46-48.
void twogrid_bigph(ssn_ctx* c, const CsrView& A, const double* b, const AmgOptions& o_in, double* x, int* it_out,
                   double* rel_res_out, double* rel_resk, double* rhok, int* hist_len, bool generic) {
    AmgOptions o = o_in;
    if (generic) {
        o.theta = 0.25; o.inter = 1;                                          // twogrid.m:73,84-88
        if (o.bigph) SSN_REQUIRE(o.fnode > 0, SSN_E_BIGPH_FNODE, "bigph = 1 requires fnode > 0");      // :36-38
    } else {
        o.bigph = 1;
    }
    if (o.bigph) SSN_REQUIRE(o.fnode > 0 && o.fnode < A.nrows, SSN_E_BIGPH_FNODE, "twogrid_bigph requires 0 < amg_options.fnode < N");
    amg_setup(c, A, o, 2);
    Phase ph_solve(c, "twogrid solve loop total");
    Hierarchy& H = *c->hier;
    SSN_REQUIRE(H.J == 2, SSN_E_INVALID, "twogrid_bigph: the system is too small for a coarse level");
    Level& L = H.lv[0]; Level& Lc = H.lv[1];
    const int n = L.N;
    if (o.guess) SSN_CUDA(cudaMemcpyAsync(x, o.guess, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    else fill_double(c, x, n, 0.0);
    ssn_pcg_options po{};
    po.retol = -1.0; po.maxit = 100; po.precd = 2; po.nf = 0; po.guess_dev = nullptr;       // twogrid_bigph.m:98
    int it = 0;
    double rel_res = 0.0;
    std::vector<double> relk(1, 1.0), rho(1, NAN);
    // ---- the whole iteration loop in ONE kernel (amg_cluster.cu: one 16-CTA cluster, level vectors in distributed shared
    // memory, the coarse PCG inside the kernel: ~100 PCG iterations per two-grid iteration at a few microseconds each instead of
    // the grid-wide pcg_kernel's 16) when both levels fit; the kernel leaves x untouched when it refuses the level-1 matrix
    if (c->cluster_solve && c->dsm_solve && c->tg_cluster && L.A.nnz + Lc.A.nnz <= c->cluster_max_nnz) {
        const int hl = o.maxit + 2;
        Buf<double> hist(c, (size_t)2 * hl);
        Buf<int> iout(c, 4);
        if (dsm_cluster_solve(c, H, b, x, o, false, hist.p, hl, iout.p, &po)) {
            int hi[4];
            read_back(c, iout.p, hi, 4);
            if (hi[2] == 0) {
                it = hi[0];
                const int len = hi[1];
                std::vector<double> hh((size_t)2 * hl);
                read_back(c, hist.p, hh.data(), hh.size());
                relk.assign(hh.begin(), hh.begin() + len);
                rho.assign(hh.begin() + hl, hh.begin() + hl + len);
                rel_res = (len > 1) ? relk[len - 1] : 0.0;
                if (it_out) *it_out = it;
                if (rel_res_out) *rel_res_out = rel_res;
                if (hist_len) *hist_len = (int)relk.size();
                if (rel_resk) std::memcpy(rel_resk, relk.data(), sizeof(double) * relk.size());
                if (rhok) std::memcpy(rhok, rho.data(), sizeof(double) * rho.size());
                amg_clear(c);
                return;
            }
        }
    }
    double h[2];
    int np = launch_resid(c, L, b, x, L.r, H.part);                          // r = b - A*x ; res0 = norm(A*x-b)   :61
    SSN_LAUNCH(c, reduce_parts_kernel, 1, 256, 0, H.part.p, np, H.scal.p);
    read_back(c, H.scal.p, h, 2);
    const double res0 = std::sqrt(h[1]);
    double res_prev = res0;
    if (res0 == 0.0) {
        rel_res = 0.0; relk.assign(1, 0.0); rho.assign(1, INFINITY);
    } else {
        it = 1;
        while (relk[it - 1] > o.retol && it <= o.maxit) {                     // :65
            smooth_host(c, H, 0, o.isnsp, 0, true);                           // :82-90
            launch_resid(c, L, L.r, (H.smoth == 0) ? nullptr : L.e.p, L.g, H.part);
            spmv(c, Lc.Pt, L.g, Lc.r);                                        // rrc = Pro'*(r-A*e)   :92
            pcg_solve(c, Lc.A, Lc.r, &po, Lc.e, nullptr, nullptr, nullptr);   // :99
            spmv_add(c, Lc.P, Lc.e, L.e);                                     // :107
            smooth_host(c, H, 0, o.isnsp, 1, false);                          // :109-116
            SSN_LAUNCH(c, axpy_kernel, cdiv(n, 256), 256, 0, n, 1.0, L.e.p, x);
            np = launch_resid(c, L, b, x, L.r, H.part);
            SSN_LAUNCH(c, reduce_parts_kernel, 1, 256, 0, H.part.p, np, H.scal.p);
            read_back(c, H.scal.p, h, 2);
            const double res = std::sqrt(h[1]);
            rel_res = res / res0;
            relk.push_back(rel_res);
            rho.push_back(res / res_prev);                                    // res/norm(r)   :71
            res_prev = res;
            ++it;
            if (rho[it - 1] > 1.0) break;                                     // :72
        }
        relk.resize(it); rho.resize(it); --it;
    }
    if (it_out) *it_out = it;
    if (rel_res_out) *rel_res_out = rel_res;
    if (hist_len) *hist_len = (int)relk.size();
    if (rel_resk) std::memcpy(rel_resk, relk.data(), sizeof(double) * relk.size());
    if (rhok) std::memcpy(rhok, rho.data(), sizeof(double) * rho.size());
    amg_clear(c);
}

// PCG.m:18-105
void pcg_solve(ssn_ctx* c, const CsrView& H, const double* e, const ssn_pcg_options* opts, double* d, int* it_out,
               double* res_out, double* resk_host) {
    SSN_REQUIRE(H.nrows == H.ncols, SSN_E_NOT_SQUARE, "PCG: matrix must be square");
    const int n = H.nrows;
    double retol = 1e-11; int maxit = 10000, precd = 2, nf = 0; const double* guess = nullptr;
    if (opts) {
        if (!(opts->retol < 0) && opts->retol == opts->retol) retol = opts->retol;
        if (opts->maxit >= 0) maxit = opts->maxit;
        if (opts->precd > 0) precd = opts->precd;
        nf = opts->nf; guess = opts->guess_dev;
    }
    SSN_REQUIRE(precd >= 1 && precd <= 5, SSN_E_UNSUPPORTED, "PCG: precd must be 1..5");
    if (precd == 5) SSN_REQUIRE(nf > 0 && nf < n, SSN_E_PCG_NF, "SSOR for bigraph requires pcg_options.nf!!!");
    Buf<double> r(c, n), p(c, n), q(c, n), w(c, n), aux(c, n), diag(c, n), resk(c, maxit > 0 ? maxit : 1), scal(c, 2);
    Buf<int> itd(c, 1);
    extract_diag(c, H, diag);
    resk.zero();
    PcgArgs a{};
    a.n = n; a.ptr = H.ptr; a.idx = H.idx; a.val = H.val; a.rhs = e; a.guess = guess;
    a.d = d; a.r = r; a.p = p; a.q = q; a.w = w; a.a = aux; a.diag = diag;
    a.precd = precd; a.nf = nf; a.maxit = maxit; a.tol2 = retol * retol;
    a.resk = resk; a.it_out = itd; a.scal_out = scal;
    // precd 3 (SSOR, PCG.m:39-44,96-99) and 4 (ichol, :45-51,100-101): the triangular factors and their
    // dependency levels are built once per call on the host (one pass over the pattern; the incomplete
    // factorisation is inherently sequential), the solves of every iteration run in the kernel.
    Buf<int> lp_d, li_d, up_d, ui_d, lrows_d, llev_d, urows_d, ulev_d;
    Buf<double> lv_d, uv_d, mid_d;
    TriFactors F;
    if ((precd == 3 || precd == 4) && c->device_setup) {
        // SSN_DEVICE_SETUP=1: the same factors, levels and row groups built on the device (trifactor.cu)
        build_tri_factors_device(c, H, precd, F);
        a.lp = F.lp; a.li = F.li; a.lv = F.lv; a.up = F.up; a.ui = F.ui; a.uv = F.uv; a.mid = F.has_mid ? F.mid.p : nullptr;
        a.lrows = F.lrows; a.llev = F.llev; a.nlev_l = F.nl; a.urows = F.urows; a.ulev = F.ulev; a.nlev_u = F.nu;
    } else if (precd == 3 || precd == 4) {
        const int64_t nnz = H.nnz;
        std::vector<int> hp((size_t)n + 1), hi((size_t)nnz); std::vector<double> hv((size_t)nnz);
        SSN_CUDA(cudaMemcpyAsync(hp.data(), H.ptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaMemcpyAsync(hi.data(), H.idx, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaMemcpyAsync(hv.data(), H.val, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<int> lp(1, 0), li, up(1, 0), ui; std::vector<double> lv, uv, mid;
        if (precd == 3) {
            const double om = 1.5, sc = om * (2.0 - om);
            mid.resize(n);
            for (int i = 0; i < n; ++i) {
                double dii = 0.0;
                for (int e = hp[i]; e < hp[i + 1]; ++e) if (hi[e] == i) dii = hv[e];
                mid[i] = dii;                                                   // p2 = D*p1            :97
                for (int e = hp[i]; e < hp[i + 1]; ++e) if (hi[e] < i) { li.push_back(hi[e]); lv.push_back(om * hv[e]); }
                li.push_back(i); lv.push_back(dii);                             // D + w*L              :96
                ui.push_back(i); uv.push_back(sc * dii);                        // (w*(2-w))*(D + w*U)  :99 (as MATLAB parses it)
                for (int e = hp[i]; e < hp[i + 1]; ++e) if (hi[e] > i) { ui.push_back(hi[e]); uv.push_back(sc * (om * hv[e])); }
                lp.push_back((int)li.size()); up.push_back((int)ui.size());
            }
        } else {
            // IC(0): L has the pattern of tril(H); row i: L_ij = (a_ij - sum_{t<j} L_it L_jt)/L_jj, L_ii = sqrt(a_ii - sum L_it^2)
            for (int i = 0; i < n; ++i) {
                const int r0 = (int)li.size();
                bool has_diag = false;
                for (int e = hp[i]; e < hp[i + 1]; ++e) {
                    const int j = hi[e];
                    if (j > i) continue;
                    double sacc = hv[e];
                    int a0 = r0, a1 = (int)li.size();                           // row i so far (columns < j)
                    int b0 = (j < i) ? lp[j] : r0, b1 = (j < i) ? lp[j + 1] - 1 : a1;   // row j without its diagonal
                    while (a0 < a1 && b0 < b1) {
                        if (li[a0] == li[b0]) { sacc -= lv[a0] * lv[b0]; ++a0; ++b0; }
                        else if (li[a0] < li[b0]) ++a0; else ++b0;
                    }
                    if (j < i) { li.push_back(j); lv.push_back(sacc / lv[lp[j + 1] - 1]); }
                    else { SSN_REQUIRE(sacc > 0.0, SSN_E_NOT_SPD, "ichol: encountered nonpositive pivot"); li.push_back(i); lv.push_back(std::sqrt(sacc)); has_diag = true; }
                }
                SSN_REQUIRE(has_diag, SSN_E_NOT_SPD, "ichol: zero on the diagonal");
                lp.push_back((int)li.size());
            }
            // Uf = L'
            std::vector<int> cnt((size_t)n + 1, 0);
            for (int v : li) ++cnt[(size_t)v + 1];
            for (int i = 0; i < n; ++i) cnt[(size_t)i + 1] += cnt[i];
            up.assign(cnt.begin(), cnt.end());
            ui.resize(li.size()); uv.resize(li.size());
            std::vector<int> pos(cnt.begin(), cnt.end() - 1);
            for (int i = 0; i < n; ++i) for (int e = lp[i]; e < lp[i + 1]; ++e) { const int j = li[e]; ui[pos[j]] = i; uv[pos[j]] = lv[e]; ++pos[j]; }
        }
        // dependency levels: forward (columns < row), backward (columns > row)
        std::vector<int> levl(n, 0), levu(n, 0);
        int nl = 0, nu = 0;
        for (int i = 0; i < n; ++i) { int l = 0; for (int e = lp[i]; e < lp[i + 1] - 1; ++e) l = std::max(l, levl[li[e]] + 1); levl[i] = l; nl = std::max(nl, l + 1); }
        for (int i = n - 1; i >= 0; --i) { int l = 0; for (int e = up[i] + 1; e < up[i + 1]; ++e) l = std::max(l, levu[ui[e]] + 1); levu[i] = l; nu = std::max(nu, l + 1); }
        auto bucket = [&](const std::vector<int>& lev, int nlev, std::vector<int>& rows, std::vector<int>& ptr) {
            ptr.assign((size_t)nlev + 1, 0);
            for (int i = 0; i < n; ++i) ++ptr[(size_t)lev[i] + 1];
            for (int l = 0; l < nlev; ++l) ptr[(size_t)l + 1] += ptr[l];
            rows.resize(n);
            std::vector<int> pos(ptr.begin(), ptr.end() - 1);
            for (int i = 0; i < n; ++i) rows[pos[lev[i]]++] = i;
        };
        std::vector<int> lrows, llev, urows, ulev;
        bucket(levl, nl, lrows, llev); bucket(levu, nu, urows, ulev);
        auto up_i = [&](Buf<int>& d, const std::vector<int>& h) { d.alloc(c, std::max<size_t>(h.size(), 1)); if (!h.empty()) SSN_CUDA(cudaMemcpyAsync(d.p, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, c->stream)); };
        auto up_d2 = [&](Buf<double>& d, const std::vector<double>& h) { d.alloc(c, std::max<size_t>(h.size(), 1)); if (!h.empty()) SSN_CUDA(cudaMemcpyAsync(d.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, c->stream)); };
        up_i(lp_d, lp); up_i(li_d, li); up_d2(lv_d, lv); up_i(up_d, up); up_i(ui_d, ui); up_d2(uv_d, uv);
        up_i(lrows_d, lrows); up_i(llev_d, llev); up_i(urows_d, urows); up_i(ulev_d, ulev);
        if (!mid.empty()) up_d2(mid_d, mid);
        SSN_CUDA(cudaStreamSynchronize(c->stream));          // the host vectors go out of scope
        a.lp = lp_d; a.li = li_d; a.lv = lv_d; a.up = up_d; a.ui = ui_d; a.uv = uv_d; a.mid = mid.empty() ? nullptr : mid_d.p;
        a.lrows = lrows_d; a.llev = llev_d; a.nlev_l = nl; a.urows = urows_d; a.ulev = ulev_d; a.nlev_u = nu;
    }
    int per_sm = 0;
    SSN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel, 256, 0));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = c->num_sms * per_sm;
    const int need = cdiv((int64_t)n * 32, 256);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    Buf<double> part(c, (size_t)3 * grid);
    a.part = part;
    void* args[] = {&a};
    SSN_CUDA(cudaLaunchCooperativeKernel((void*)pcg_kernel, dim3(grid), dim3(256), args, 0, c->stream));
    c->launches++;
    int it = read_scalar(c, itd.p);
    double s[2]; read_back(c, scal.p, s, 2);
    if (it_out) *it_out = it;
    if (res_out) *res_out = std::sqrt(std::fabs(s[0] / s[1]));             // PCG.m:87 (0/0 -> NaN)
    if (resk_host && maxit > 0) read_back(c, resk.p, resk_host, (size_t)maxit);
}

}  // namespace ssn
