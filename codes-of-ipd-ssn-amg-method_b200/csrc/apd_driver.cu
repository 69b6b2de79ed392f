// apd_driver.cu -- the reference's Class 1 script body as ONE entry point:
// warm start (Class1/warmup_class1.m), the APD outer loop and the semismooth-Newton inner loop with its Armijo
// line search and KKT bookkeeping (Class1/APD_SsN_Class1.m:32-275), every plan-sized array resident on the
// device, no interpreter between the kernels.  The plan-wide lines run through the fused kernels of plan_ops.cu
// (ssn_apd_begin / ssn_apd_end / ssn_prox_residual / ssn_linesearch), the (n+m)-sized lines through the small
// kernels below, the inner linear solve through the selected solver (2 PCG on Jk, 3 aug_PCG, 4 Hybrid_AMG,
// 5 Hybrid_twogrid).  The host reads back only what the loop control of the script needs.
#include "amg.cuh"
#include "plan_ops.cuh"
#include "solvers.cuh"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>

namespace ssn {

namespace {

constexpr int kVT = 1024;

// out[0..3] = { sum a1*b1, sum a2*b2, sum a3*b3, sum a4*b4 } over N entries; null pairs give 0; b == nullptr: a*a.
// One block, fixed order (deterministic); N = n+m is small.
__global__ void __launch_bounds__(kVT) dots4_kernel(int64_t N, const double* a1, const double* b1, const double* a2, const double* b2,
                                                     const double* a3, const double* b3, const double* a4, const double* b4, double* out) {
    __shared__ double red[32];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = threadIdx.x; i < N; i += kVT) {
        if (a1) { const double x = a1[i]; s[0] = fma(x, b1 ? b1[i] : x, s[0]); }
        if (a2) { const double x = a2[i]; s[1] = fma(x, b2 ? b2[i] : x, s[1]); }
        if (a3) { const double x = a3[i]; s[2] = fma(x, b3 ? b3[i] : x, s[2]); }
        if (a4) { const double x = a4[i]; s[3] = fma(x, b4 ? b4[i] : x, s[3]); }
    }
    for (int k = 0; k < 4; ++k) { const double t = block_sum(s[k], red); if (threadIdx.x == 0) out[k] = t; __syncthreads(); }
}

// wlk = bk1*(lk - 1/bk*(axk - b)) - b                                   APD_SsN_Class1.m:126
__global__ void wlk_kernel(int64_t N, const double* lk, const double* axk, const double* b, double bk1, double inv_bk, double* wlk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double t = __dsub_rn(lk[i], __dmul_rn(inv_bk, __dsub_rn(axk[i], b[i])));
    wlk[i] = __dsub_rn(__dmul_rn(bk1, t), b[i]);
}

// Fk = bk1*lk - Axprox - wlk (:130,:144,:212);  mFk = -Fk (the right-hand side prob_data.z, :156)
__global__ void fk_kernel(int64_t N, const double* lk, const double* axp, const double* wlk, double bk1, double* Fk, double* mFk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double f = __dsub_rn(__dsub_rn(__dmul_rn(bk1, lk[i]), axp[i]), wlk[i]);
    Fk[i] = f;
    if (mFk) mFk[i] = -f;
}

// d = a - b (plan-sized helper of the KKT residual at the start / after a restart)
__global__ void sub_kernel(int64_t n, const double* a, const double* b, double* d) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = a[i] - b[i];
}

// Class 2 line-search trial (Class2/APD_SsN_Class2.m:199-201, 205-207): lk_new = lk_old + alpha*zeta and
// out2 = { ||lk_new||^2, wlk'*lk_new }; one block, fixed order
__global__ void __launch_bounds__(kVT) pot_trial_kernel(int64_t N1, const double* lk_old, const double* zeta, double alpha, const double* wlk,
                                                         double* lk_new, double* out2) {
    __shared__ double red[32];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = threadIdx.x; i < N1; i += kVT) {
        const double v = __dadd_rn(lk_old[i], __dmul_rn(alpha, zeta[i]));
        lk_new[i] = v;
        s0 = fma(v, v, s0); s1 = fma(wlk[i], v, s1);
    }
    s0 = block_sum(s0, red); s1 = block_sum(s1, red);
    if (threadIdx.x == 0) { out2[0] = s0; out2[1] = s1; }
}

// squared norms of three consecutive blocks of d = a - b: [0, n0), [n0, n1), [n1, n2)  (KKT residuals of x, y, z)
__global__ void __launch_bounds__(kVT) diff_norms3_kernel(const double* a, const double* b, int64_t n0, int64_t n1, int64_t n2, double* part) {
    __shared__ double red[32];
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * kVT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kVT) {
        const double d = a[i] - b[i];
        const int k = (i < n0) ? 0 : ((i < n1) ? 1 : 2);
        s[k] = fma(d, d, s[k]);
    }
    for (int k = 0; k < 3; ++k) { const double t = block_sum(s[k], red); if (threadIdx.x == 0) part[3 * blockIdx.x + k] = t; __syncthreads(); }
}
__global__ void __launch_bounds__(256) sum3_kernel(const double* part, int nb, double* out3) {
    __shared__ double red[32];
    double s[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nb; b += 256) { s[0] += part[3 * b]; s[1] += part[3 * b + 1]; s[2] += part[3 * b + 2]; }
    for (int k = 0; k < 3; ++k) { const double t = block_sum(s[k], red); if (threadIdx.x == 0) out3[k] = t; __syncthreads(); }
}
// t1 = [x - c ; s]: the argument whose prox(t1 - H'lk) is the projection of the KKT residual
__global__ void kkt_arg_kernel(int64_t mn, int64_t L, const double* u, const double* cost, double* t1) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x)
        t1[i] = (i < mn) ? (u[i] - cost[i]) : u[i];
}
// huk = [Ax(x) + s ; phi'x] - b, in place on ax (N+1 entries; ax[N] preset to phi'x)
__global__ void hub_kernel(int64_t N, const double* us, const double* b, double* ax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) ax[i] = ax[i] + us[i] - b[i];
    else if (i == N) ax[N] = ax[N] - b[N];
}
__global__ void set1_kernel(double* dst, double v) { if (threadIdx.x == 0 && blockIdx.x == 0) dst[0] = v; }

struct Clock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double s() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

}  // namespace

// ONE semismooth-Newton step of Class1/APD_SsN_Class1.m:137-212 at a fixed APD state (wk, wlk, bk1, tk) from the duals lk:
// fused residual + active set -> ASAt -> the inner solve -> Armijo line search -> new residual, with no interpreter and
// no plan-sized host traffic between them.  info (host, 12 doubles): E, nnz(H0), components, it_num, inner iterations,
// relative residual of the inner solve, ll, reads of wk by the line search, ||F(lk)||, ||F(lk_new)||, ms of the plan-wide
// part / of the inner solve (host clock; device-synchronised only when the phase profiler is on).
void ssn_step_class1(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                     int64_t n, double bk1, double tk, const double* gama, double gama_s, int inner_solver, const ssn_amg_options* amg_in,
                     double* lk_new, double* Fk_new, double* info12) {
    SSN_REQUIRE(wk && lk && wlk && p && q && lk_new && Fk_new && m > 0 && n > 0, SSN_E_INVALID, "ssn_step_class1: bad arguments");
    SSN_REQUIRE(inner_solver == 4 || inner_solver == 5, SSN_E_UNSUPPORTED, "ssn_step_class1: inner_solver must be 4 (Hybrid_AMG) or 5 (Hybrid_twogrid)");
    const int64_t N = m + n, mn = m * n;
    const double nu = 0.2, delta = 0.9; const int ll_max = 500;                                      // :36
    ssn_amg_options amg{}; amg.retol = 1e-11; amg.bigph = 1; amg.maxit = 30; amg.theta = 0.25; amg.smoth = 5; amg.cycle = 'w';
    amg.isnsp = 1; amg.inter = 1; amg.fnode = 0; amg.guess_dev = nullptr;                           // :87-88
    if (amg_in) amg = *amg_in;
    const int gN = cdiv(N, 256);
    Buf<double> axp(c, N), Fk(c, N), mFk(c, N), zeta(c, N), scal(c, 8);
    Buf<uint8_t> s(c, mn);
    double h[4];
    Clock tp0;
    plan_prox_residual(c, wk, lk, p, q, m, n, tk, gama, gama_s, axp, nullptr, nullptr, s, scal.p + 4);   // :139-144
    SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N, lk, axp.p, wlk, bk1, Fk.p, mFk.p);
    double n2_old, Ecount;
    { double hh[2]; read_back(c, scal.p + 4, hh, 2); n2_old = hh[0]; Ecount = hh[1]; }
    double plan_ms = tp0.s() * 1e3;
    Csr H0 = asat(c, s, p, q, m, n);                                                                  // :142
    Clock ts;
    ssn_csr Hv = H0.view();
    ssn_prob_data pd{}; pd.bk1 = bk1; pd.tk = tk; pd.m = m; pd.n = n; pd.p_dev = p; pd.q_dev = q; pd.t_dev = nullptr;
    pd.H0 = &Hv; pd.z_dev = mFk.p; pd.s_dev = nullptr; pd.phi_dev = nullptr;                          // :154-156
    int itl = 0, inf2[2] = {0, 0}; double resl = 0.0;
    hybrid_amg(c, &pd, &amg, zeta.p, &itl, &resl, inf2, inner_solver == 5);                           // :161 / :178
    if (c->prof) SSN_CUDA(cudaStreamSynchronize(c->stream));
    const double solve_ms = ts.s() * 1e3;
    Clock tp1;
    SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N, lk, (const double*)nullptr, wlk, lk, Fk.p, zeta.p, Fk.p, (const double*)nullptr, scal.p);   // :182, :198
    read_back(c, scal.p, h, 4);
    const double f0 = bk1 / 2 * h[0] - h[1];
    const double cFk_old = f0 + 0.5 * tk * n2_old;
    const double ress = std::fabs(h[2]);
    const double nFo = std::sqrt(h[3]);
    int ll = 0, passes = 0; double n2_new = 0.0, cF_new = 0.0;
    plan_linesearch(c, wk, lk, zeta, wlk, p, q, m, n, tk, bk1, gama, gama_s, nu, delta, ll_max, cFk_old, ress, 0, lk_new, &ll, &n2_new,
                    &cF_new, &passes);                                                                // :189-211
    plan_prox_residual(c, wk, lk_new, p, q, m, n, tk, gama, gama_s, axp, nullptr, nullptr, nullptr, scal.p + 4);   // :212
    SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N, lk_new, axp.p, wlk, bk1, Fk_new, (double*)nullptr);
    SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N, Fk_new, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr,
               (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, scal.p);
    read_back(c, scal.p, h, 1);
    plan_ms += tp1.s() * 1e3;
    if (info12) {
        info12[0] = Ecount; info12[1] = (double)H0.nnz; info12[2] = inf2[0]; info12[3] = inf2[1]; info12[4] = itl; info12[5] = resl;
        info12[6] = ll; info12[7] = passes; info12[8] = nFo; info12[9] = std::sqrt(h[0]); info12[10] = plan_ms; info12[11] = solve_ms;
    }
}

void apd_ssn_class1(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                    int64_t n, const double* gama, double gama_s, const ssn_apd_options* op, double* xk_out, double* lk_out,
                    ssn_apd_result* res, double* fxk_hist, double* kktx_hist, double* kktl_hist, int32_t* ssn_its_hist,
                    double* steps_host, int64_t steps_cap) {
    SSN_REQUIRE(cost && r && l && p && q && m > 0 && n > 0 && xk_out && lk_out, SSN_E_INVALID, "apd_ssn_class1: bad arguments");
    const int64_t N = m + n, mn = m * n;
    const int inner_solver = (op && op->inner_solver > 0) ? op->inner_solver : 4;                  // :70
    const int maxit = (op && op->maxit > 0) ? op->maxit : 100;                                      // :35
    const double KKT_Tol = (op && op->KKT_Tol > 0) ? op->KKT_Tol : 1e-6;
    const int warm_maxit = (op && op->warm_maxit >= 0) ? op->warm_maxit : 100;                      // :59
    const int max_outer = (op && op->max_outer > 0) ? op->max_outer : 0;
    const double max_seconds = (op && op->max_seconds > 0) ? op->max_seconds : 0.0;
    SSN_REQUIRE(inner_solver >= 2 && inner_solver <= 5, SSN_E_UNSUPPORTED, "inner_solver must be 2 (PCG), 3 (aug_PCG), 4 (Hybrid_AMG) or 5 (Hybrid_twogrid)");
    const int SsN_IT = 50; const double SsN_Tol1 = 1e-11, nu = 0.2, delta = 0.9; const int ll_max = 500;   // :36
    ssn_amg_options amg{}; amg.retol = 1e-11; amg.bigph = 1; amg.maxit = 30; amg.theta = 0.25; amg.smoth = 5; amg.cycle = 'w';
    amg.isnsp = 1; amg.inter = 1; amg.fnode = 0; amg.guess_dev = nullptr;                           // :87-88
    if (op && op->amg) amg = *op->amg;
    ssn_pcg_options pcg{}; pcg.retol = 1e-11; pcg.maxit = 10000; pcg.precd = 2; pcg.nf = 0; pcg.guess_dev = nullptr;   // :81
    if (op && op->pcg) pcg = *op->pcg;
    if (inner_solver == 2 && pcg.precd == 5) pcg.nf = (int)n;

    const int gN = cdiv(N, 256);
    Buf<double> b(c, N), xk(c, mn), vk(c, mn), xk1(c, mn), vk1(c, mn), wk(c, mn);
    Buf<double> lk(c, N), lk_new(c, N), lk_old(c, N), wlk(c, N), axk(c, N), axp(c, N), Fk(c, N), mFk(c, N), zeta(c, N), axk1(c, N);
    Buf<double> scal(c, 8);
    Buf<uint8_t> s(c, mn);
    SSN_CUDA(cudaMemcpyAsync(b.p, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));       // b = [r; l]   :33
    SSN_CUDA(cudaMemcpyAsync(b.p + n, l, sizeof(double) * m, cudaMemcpyDeviceToDevice, c->stream));
    double h[8];
    auto dots = [&](const double* a1, const double* b1, const double* a2, const double* b2, const double* a3, const double* b3,
                    const double* a4, const double* b4) {
        SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N, a1, b1, a2, b2, a3, b3, a4, b4, scal.p);
        read_back(c, scal.p, h, 4);
    };
    // KKT residuals of (x, lam): || x - prox(x - c - A'lam) ||, || Ax - b ||     :62-65, :252-254
    auto kkt = [&](const double* x, const double* lam, double& kx, double& kl, double& cx) {
        Buf<double> t1(c, mn), t2(c, mn), ax(c, N);
        SSN_LAUNCH(c, sub_kernel, 2048, 256, 0, mn, x, cost, t1.p);
        plan_prox_residual(c, t1, lam, p, q, m, n, 1.0, gama, gama_s, nullptr, t2, nullptr, nullptr, scal.p + 4);
        SSN_LAUNCH(c, sub_kernel, 2048, 256, 0, mn, x, t2.p, t1.p);
        kx = std::sqrt(dev_dot(c, t1, t1, mn));
        cx = dev_dot(c, cost, x, mn);
        plan_ax(c, x, p, q, m, n, ax);
        SSN_LAUNCH(c, sub_kernel, 64, 256, 0, N, ax.p, b.p, ax.p);
        dots(ax.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        kl = std::sqrt(h[0]);
    };

    Clock t_all;
    plan_warmup_class1(c, cost, b, p, q, m, n, gama, gama_s, warm_maxit, xk, lk);                     // :59
    SSN_CUDA(cudaMemcpyAsync(vk.p, xk.p, sizeof(double) * mn, cudaMemcpyDeviceToDevice, c->stream));  // vk = xk
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    const double t_warm = t_all.s();
    double kx0, kl0, cx0;
    kkt(xk, lk, kx0, kl0, cx0);
    std::vector<double> fx(1, cx0), KX(1, kx0), KL(1, kl0);
    double bk = 1.0;
    int ssn_total = 0, ls_trials = 0, ls_passes = 0, amg_calls = 0, converged = 0;
    int64_t nsteps = 0;
    double solve_s = 0.0, asat_s = 0.0, plan_s = 0.0;
    double rr0 = INFINITY, rr1 = INFINITY;
    Clock t_loop;
    int k = 0;
    for (k = 1; k <= maxit; ++k) {                                                                   // :101
        const double resk = std::max(KX[k - 1], KL[k - 1]);
        const double ak = std::sqrt((double)k * (double)k * bk);                                      // :113
        double bk1 = bk / (1 + ak); const double tk = bk * (1 + ak) / (ak * ak);                      // :120
        const double SsN_Tol = std::max(bk1 / ((double)k * (double)k), SsN_Tol1);                     // :123
        plan_apd_begin(c, cost, xk, vk, p, q, m, n, ak, bk, wk, axk);                                 // :125 + Ax(xk) of :126
        SSN_LAUNCH(c, wlk_kernel, gN, 256, 0, N, lk.p, axk.p, b.p, bk1, 1 / bk, wlk.p);               // :126
        SSN_CUDA(cudaMemcpyAsync(lk_new.p, lk.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, c->stream));
        plan_prox_residual(c, wk, lk_new, p, q, m, n, tk, gama, gama_s, axp, nullptr, nullptr, s, scal.p + 4);   // :129-130 (+ s for :140)
        SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N, lk_new.p, axp.p, wlk.p, bk1, Fk.p, mFk.p);
        dots(Fk.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        double nF = std::sqrt(h[0]);
        double n2_old, Ecount;
        { double hh[2]; read_back(c, scal.p + 4, hh, 2); n2_old = hh[0]; Ecount = hh[1]; }
        double Fk_res = nF;
        int ssn_it = 0;
        while (nF > SsN_Tol) {                                                                       // :137
            ++ssn_it;
            std::swap(lk_old.p, lk_new.p);                                                            // lk_old = lk_new; z, s, prox of lk_old are evaluated already
            Clock ta;
            Csr H0 = asat(c, s, p, q, m, n);                                                          // :142
            if (c->prof) SSN_CUDA(cudaStreamSynchronize(c->stream));
            asat_s += ta.s();
            Clock ts;
            ssn_csr Hv = H0.view();
            ssn_prob_data pd{}; pd.bk1 = bk1; pd.tk = tk; pd.m = m; pd.n = n; pd.p_dev = p; pd.q_dev = q; pd.t_dev = nullptr;
            pd.H0 = &Hv; pd.z_dev = mFk.p; pd.s_dev = nullptr; pd.phi_dev = nullptr;                  // :154-156
            int itl = 0, info[2] = {0, 0}; double resl = 0.0;
            if (inner_solver == 2) {                                                                  // :149-152
                Csr Jk; jk_system(c, &pd, Jk);
                pcg_solve(c, Jk, mFk.p, &pcg, zeta.p, &itl, &resl, nullptr);
            } else if (inner_solver == 3) {
                aug_pcg(c, &pd, &pcg, zeta.p, &itl, &resl, info);                                     // :158
            } else {
                hybrid_amg(c, &pd, &amg, zeta.p, &itl, &resl, info, inner_solver == 5);               // :161 / :178
                ++amg_calls;
            }
            solve_s += ts.s();
            Clock tp;
            dots(lk_old.p, nullptr, wlk.p, lk_old.p, Fk.p, zeta.p, Fk.p, nullptr);                     // :182, :198
            const double f0 = bk1 / 2 * h[0] - h[1];
            const double cFk_old = f0 + 0.5 * tk * n2_old;
            const double ress = std::fabs(h[2]);
            const double nFo = std::sqrt(h[3]);
            int ll = 0, passes = 0; double n2_new = 0.0, cF_new = 0.0;
            plan_linesearch(c, wk, lk_old, zeta, wlk, p, q, m, n, tk, bk1, gama, gama_s, nu, delta, ll_max, cFk_old, ress, 0,
                            lk_new, &ll, &n2_new, &cF_new, &passes);                                  // :189-211
            ls_trials += ll + 1; ls_passes += passes;
            const double E_step = Ecount;
            plan_prox_residual(c, wk, lk_new, p, q, m, n, tk, gama, gama_s, axp, nullptr, nullptr, s, scal.p + 4);   // :212 (+ next s)
            SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N, lk_new.p, axp.p, wlk.p, bk1, Fk.p, mFk.p);
            dots(Fk.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            nF = std::sqrt(h[0]);
            { double hh[2]; read_back(c, scal.p + 4, hh, 2); n2_old = hh[0]; Ecount = hh[1]; }
            plan_s += tp.s();
            if (steps_host && nsteps < steps_cap) {
                double* st = steps_host + 7 * nsteps;
                st[0] = k; st[1] = ssn_it; st[2] = E_step; st[3] = info[0]; st[4] = itl; st[5] = ll; st[6] = nF;
            }
            ++nsteps;
            if (op && op->verbose) fprintf(stderr, "   SsN: it=%3d |Fk|=%.2e ll=%3d info=[%d, %d] its=%d res=%.2e E=%.0f\n", ssn_it, nF, ll, info[0], info[1], itl, resl, Ecount);
            if (nF <= SsN_Tol) break;
            if (std::fabs(nFo - nF) < SsN_Tol / 100) break;                                           // :219
            if (ssn_it == SsN_IT) break;
            if (Fk_res / nF >= 2) Fk_res = nF;
        }
        ssn_total += ssn_it;
        // :239-254 in one pass: xk1 = prox(zk), vk1, Ax(xk1), c'xk1 and the KKT residual of xk1
        plan_apd_end(c, cost, wk, xk, lk_new, p, q, m, n, tk, ak, gama, gama_s, xk1, vk1, axk1, scal.p + 4);
        SSN_LAUNCH(c, sub_kernel, 64, 256, 0, N, axk1.p, b.p, axk1.p);
        dots(axk1.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        double kl = std::sqrt(h[0]), cx, kx;
        { double hh[2]; read_back(c, scal.p + 4, hh, 2); cx = hh[0]; kx = std::sqrt(hh[1]); }
        rr0 = kx / (1 + KX[0]); rr1 = kl / (1 + KL[0]);
        if (bk1 < 1e-8 && std::max(rr0, rr1) > resk) {                                               // :245-249: restart
            // xk1 = xk; lk1 = lk; vk1 = xk; bk1 = rand
            Buf<double> rnd(c, 1);
            rng_rand(c, 1, rnd);
            bk1 = read_scalar(c, rnd.p);
            SSN_CUDA(cudaMemcpyAsync(vk.p, xk.p, sizeof(double) * mn, cudaMemcpyDeviceToDevice, c->stream));
            kkt(xk, lk, kx, kl, cx);
        } else {
            std::swap(xk.p, xk1.p); std::swap(vk.p, vk1.p); std::swap(lk.p, lk_new.p);               // :251
        }
        bk = bk1;
        fx.push_back(cx); KL.push_back(kl); KX.push_back(kx);
        if (ssn_its_hist) ssn_its_hist[k - 1] = ssn_it;
        rr0 = KX[k] / (1 + KX[0]); rr1 = KL[k] / (1 + KL[0]);
        if (op && op->verbose) fprintf(stderr, "APD: it=%3d KKT(xk)=%.2e KKT(lk)=%.2e fk=%.8e t=%.2fs\n", k, rr0, rr1, cx, t_loop.s());
        if (std::max(rr0, rr1) <= KKT_Tol) { converged = 1; break; }                                  // :266
        if (max_outer > 0 && k >= max_outer) break;
        if (max_seconds > 0 && t_loop.s() > max_seconds) break;
    }
    if (k > maxit) k = maxit;
    SSN_CUDA(cudaMemcpyAsync(xk_out, xk.p, sizeof(double) * mn, cudaMemcpyDeviceToDevice, c->stream));
    SSN_CUDA(cudaMemcpyAsync(lk_out, lk.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, c->stream));
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    if (res) {
        res->outer_its = k; res->converged = converged; res->rel_kkt = std::max(rr0, rr1); res->objective = fx.back();
        res->ssn_steps = ssn_total; res->ls_trials = ls_trials; res->ls_passes = ls_passes; res->amg_calls = amg_calls;
        res->warmup_s = t_warm; res->loop_s = t_loop.s(); res->solve_s = solve_s; res->asat_s = asat_s; res->plan_s = plan_s;
        res->hist_len = (int)fx.size(); res->steps_len = nsteps;
    }
    for (size_t i = 0; i < fx.size(); ++i) {
        if (fxk_hist) fxk_hist[i] = fx[i];
        if (kktx_hist) kktx_hist[i] = KX[i];
        if (kktl_hist) kktl_hist[i] = KL[i];
    }
}

// ===================================================================================== Class 2 (partial OT)

namespace {

// the Armijo loop of Class2/APD_SsN_Class2.m:196-213: one fused pass over wk and phi per trial, one host read per trial
// (the two dots of f0 and the norm).  Leaves lk_new, H*prox(zk) at lk_new in hp and (when asked) the next s / t.
int pot_linesearch(ssn_ctx* c, const double* wk, const double* lk_old, const double* zeta, const double* wlk, const double* p, const double* q,
                   int64_t m, int64_t n, double tk, double bk1, const double* phi, double nu, double delta, int ll_max, double cFk_old,
                   double ress, double* lk_new, double* hp, uint8_t* s_out, double* t_out, double* scal /* >= 5 doubles */, double* n2_out,
                   double* count_out) {
    const int64_t N1 = m + n + 1;
    int ll = 0;
    double h5[5];
    while (true) {
        const double alpha = std::pow(delta, (double)ll);
        SSN_LAUNCH(c, pot_trial_kernel, 1, kVT, 0, N1, lk_old, zeta, alpha, wlk, lk_new, scal);
        plan_prox_residual_pot(c, wk, lk_new, p, q, m, n, tk, phi, hp, nullptr, s_out, t_out, scal + 2);
        read_back(c, scal, h5, 5);
        const double f0 = bk1 / 2 * h5[0] - h5[1];                                                    // :200, :206
        const double cFk_new = f0 + 0.5 * tk * h5[2];
        if (!(cFk_new > cFk_old - nu * alpha * ress) || ll == ll_max) break;                          // :204, :209
        ++ll;
    }
    if (n2_out) *n2_out = h5[2];
    if (count_out) *count_out = h5[3];
    return ll;
}

}  // namespace

// ONE semismooth-Newton step of Class2/APD_SsN_Class2.m:137-217 at a fixed APD state (wk of m*n + N entries, wlk of N+1,
// bk1, tk) from the duals lk (N+1): fused residual + active flags -> ASAt -> AMG4POT (inner_solver 4; 5: its 'twogrid'
// variant; 3: PCG4POT) -> the Armijo loop -> the new residual.  info12 (host) as ssn_step_class1.
void ssn_step_class2(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                     int64_t n, double bk1, double tk, const double* phi, int inner_solver, const ssn_amg_options* amg_in,
                     const ssn_pcg_options* pcg_in, double* lk_new, double* Fk_new, double* info12) {
    SSN_REQUIRE(wk && lk && wlk && p && q && phi && lk_new && Fk_new && m > 0 && n > 0, SSN_E_INVALID, "ssn_step_class2: bad arguments");
    SSN_REQUIRE(inner_solver >= 3 && inner_solver <= 5, SSN_E_UNSUPPORTED, "ssn_step_class2: inner_solver must be 3 (PCG4POT), 4 (AMG4POT) or 5 (AMG4POT, twogrid)");
    const int64_t N = m + n, N1 = N + 1, mn = m * n;
    const double nu = 0.2, delta = 0.9; const int ll_max = 500;                                       // :28
    ssn_amg_options amg{}; amg.retol = 1e-11; amg.bigph = 1; amg.maxit = 40; amg.theta = 0.25; amg.smoth = 10; amg.cycle = 'w';
    amg.isnsp = 1; amg.inter = 1; amg.fnode = 0; amg.guess_dev = nullptr;                             // :80-81
    if (amg_in) amg = *amg_in;
    ssn_pcg_options pcg{}; pcg.retol = 1e-11; pcg.maxit = 10000; pcg.precd = 2; pcg.nf = 0; pcg.guess_dev = nullptr;   // :74
    if (pcg_in) pcg = *pcg_in;
    const int gN = cdiv(N1, 256);
    Buf<double> hp(c, N1), Fk(c, N1), mFk(c, N1), zeta(c, N1), tt(c, N), scal(c, 8);
    Buf<uint8_t> s(c, mn);
    double h[5];
    Clock tp0;
    plan_prox_residual_pot(c, wk, lk, p, q, m, n, tk, phi, hp, nullptr, s, tt, scal.p + 2);          // :139-150
    SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N1, lk, hp.p, wlk, bk1, Fk.p, mFk.p);
    double n2_old, Ecount;
    { double hh[2]; read_back(c, scal.p + 2, hh, 2); n2_old = hh[0]; Ecount = hh[1]; }
    double plan_ms = tp0.s() * 1e3;
    Csr H0 = asat(c, s, p, q, m, n);                                                                  // :146
    Clock ts;
    ssn_csr Hv = H0.view();
    ssn_prob_data pd{}; pd.bk1 = bk1; pd.tk = tk; pd.m = m; pd.n = n; pd.p_dev = p; pd.q_dev = q; pd.t_dev = tt.p;
    pd.H0 = &Hv; pd.z_dev = mFk.p; pd.s_dev = s.p; pd.phi_dev = phi;                                  // :163-166
    int itl = 0, inf2[2] = {0, 0}; double resl = 0.0;
    if (inner_solver == 3) pcg4pot(c, &pd, &pcg, zeta.p, &itl, &resl, inf2);                          // :168
    else amg4pot(c, &pd, &amg, zeta.p, &itl, &resl, inf2, inner_solver == 5);                         // :171 / :182
    if (c->prof) SSN_CUDA(cudaStreamSynchronize(c->stream));
    const double solve_ms = ts.s() * 1e3;
    Clock tp1;
    SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N1, lk, (const double*)nullptr, wlk, lk, Fk.p, zeta.p, Fk.p, (const double*)nullptr, scal.p);   // :196-197, :203
    read_back(c, scal.p, h, 4);
    const double cFk_old = bk1 / 2 * h[0] - h[1] + 0.5 * tk * n2_old;
    const double ress = std::fabs(h[2]);
    const double nFo = std::sqrt(h[3]);
    const int ll = pot_linesearch(c, wk, lk, zeta, wlk, p, q, m, n, tk, bk1, phi, nu, delta, ll_max, cFk_old, ress, lk_new, hp, nullptr, nullptr,
                                  scal, nullptr, nullptr);
    SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N1, lk_new, hp.p, wlk, bk1, Fk_new, (double*)nullptr);       // :217
    SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N1, Fk_new, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr,
               (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, scal.p);
    read_back(c, scal.p, h, 1);
    plan_ms += tp1.s() * 1e3;
    if (info12) {
        info12[0] = Ecount; info12[1] = (double)H0.nnz; info12[2] = inf2[0]; info12[3] = inf2[1]; info12[4] = itl; info12[5] = resl;
        info12[6] = ll; info12[7] = ll + 1; info12[8] = nFo; info12[9] = std::sqrt(h[0]); info12[10] = plan_ms; info12[11] = solve_ms;
    }
}

// The Class 2 script as one call: warm start (Class2/warmup_class2.m), APD outer loop and SsN inner loop with its Armijo
// line search and KKT bookkeeping (Class2/APD_SsN_Class2.m:25-285), every (m*n + N)-sized array resident on the device.
// kkt4_hist: 4 doubles per outer iteration, {KKT_xk, KKT_yk, KKT_zk, KKT_lk} (unscaled, as the script stores them).
void apd_ssn_class2(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                    int64_t n, double mu, const double* phi, const ssn_apd_options* op, double* uk_out, double* lk_out,
                    ssn_apd_result* res, double* fxk_hist, double* kkt4_hist, int32_t* ssn_its_hist, double* steps_host, int64_t steps_cap) {
    SSN_REQUIRE(cost && r && l && p && q && phi && m > 0 && n > 0 && uk_out && lk_out, SSN_E_INVALID, "apd_ssn_class2: bad arguments");
    const int64_t N = m + n, N1 = N + 1, mn = m * n, L = mn + N;
    const int inner_solver = (op && op->inner_solver > 0) ? op->inner_solver : 4;
    const int maxit = (op && op->maxit > 0) ? op->maxit : 100;                                      // :27
    const double KKT_Tol = (op && op->KKT_Tol > 0) ? op->KKT_Tol : 1e-6;
    const int warm_maxit = (op && op->warm_maxit >= 0) ? op->warm_maxit : 100;                      // :48
    const int max_outer = (op && op->max_outer > 0) ? op->max_outer : 0;
    const double max_seconds = (op && op->max_seconds > 0) ? op->max_seconds : 0.0;
    SSN_REQUIRE(inner_solver >= 3 && inner_solver <= 5, SSN_E_UNSUPPORTED, "inner_solver must be 3 (PCG4POT), 4 (AMG4POT) or 5 (AMG4POT, twogrid)");
    const int SsN_IT = 50; const double SsN_Tol1 = 1e-10, nu = 0.2, delta = 0.9; const int ll_max = 500;   // :28
    ssn_amg_options amg{}; amg.retol = 1e-11; amg.bigph = 1; amg.maxit = 40; amg.theta = 0.25; amg.smoth = 10; amg.cycle = 'w';
    amg.isnsp = 1; amg.inter = 1; amg.fnode = 0; amg.guess_dev = nullptr;                           // :80-81
    if (op && op->amg) amg = *op->amg;
    ssn_pcg_options pcg{}; pcg.retol = 1e-11; pcg.maxit = 10000; pcg.precd = 2; pcg.nf = 0; pcg.guess_dev = nullptr;   // :74
    if (op && op->pcg) pcg = *op->pcg;

    const int gN = cdiv(N1, 256);
    Buf<double> b(c, N1), uk(c, L), vk(c, L), uk1(c, L), vk1(c, L), wk(c, L);
    Buf<double> lk(c, N1), lk_new(c, N1), lk_old(c, N1), wlk(c, N1), huk(c, N1), hp(c, N1), Fk(c, N1), mFk(c, N1), zeta(c, N1), tt(c, N);
    Buf<double> scal(c, 16);
    Buf<uint8_t> s(c, mn);
    SSN_CUDA(cudaMemcpyAsync(b.p, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));       // b = [r; l; mu]   :25
    SSN_CUDA(cudaMemcpyAsync(b.p + n, l, sizeof(double) * m, cudaMemcpyDeviceToDevice, c->stream));
    SSN_LAUNCH(c, set1_kernel, 1, 32, 0, b.p + N, mu);
    double h[8];
    // KKT residuals of (u, lam), :54-58 / :246-249 -- operator by operator (start and restarts only)
    auto kkt = [&](const double* u, const double* lam, double (&k4)[4], double& cx) {
        Buf<double> t1(c, L), t2(c, L), hx(c, N1), part(c, 3 * 256);
        SSN_LAUNCH(c, kkt_arg_kernel, 2048, 256, 0, mn, L, u, cost, t1.p);
        plan_prox_residual_pot(c, t1, lam, p, q, m, n, 1.0, phi, nullptr, t2, nullptr, nullptr, scal.p + 8);
        SSN_LAUNCH(c, diff_norms3_kernel, 256, kVT, 0, u, t2.p, mn, mn + n, L, part.p);
        SSN_LAUNCH(c, sum3_kernel, 1, 256, 0, part.p, 256, scal.p + 8);
        double h3[3]; read_back(c, scal.p + 8, h3, 3);
        k4[0] = std::sqrt(h3[0]); k4[1] = std::sqrt(h3[1]); k4[2] = std::sqrt(h3[2]);
        cx = dev_dot(c, cost, u, mn);
        plan_ax(c, u, p, q, m, n, hx);
        SSN_LAUNCH(c, set1_kernel, 1, 32, 0, hx.p + N, dev_dot(c, phi, u, mn));
        SSN_LAUNCH(c, hub_kernel, gN, 256, 0, N, u + mn, b.p, hx.p);
        SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N1, hx.p, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr,
                   (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, scal.p);
        read_back(c, scal.p, h, 1);
        k4[3] = std::sqrt(h[0]);
    };

    Clock t_all;
    plan_warmup_class2(c, cost, b, p, q, m, n, phi, warm_maxit, uk, lk);                              // :50
    SSN_CUDA(cudaMemcpyAsync(vk.p, uk.p, sizeof(double) * L, cudaMemcpyDeviceToDevice, c->stream));   // vk = uk  :51
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    const double t_warm = t_all.s();
    double K0[4], cx0;
    kkt(uk, lk, K0, cx0);
    std::vector<double> fx(1, cx0);
    std::vector<std::array<double, 4>> KK(1, {K0[0], K0[1], K0[2], K0[3]});
    double bk = 1.0;
    int ssn_total = 0, ls_trials = 0, amg_calls = 0, converged = 0;
    int64_t nsteps = 0;
    double solve_s = 0.0, asat_s = 0.0, plan_s = 0.0, rrmax = INFINITY;
    Clock t_loop;
    int k = 0;
    for (k = 1; k <= maxit; ++k) {                                                                   // :95
        const double resk = std::max(std::max(KK[k - 1][0], KK[k - 1][1]), std::max(KK[k - 1][2], KK[k - 1][3]));
        const double ak = std::sqrt((double)k * (double)k * bk);                                      // :116
        double bk1 = bk / (1 + ak); const double tk = bk * (1 + ak) / (ak * ak);
        const double SsN_Tol = std::max(bk1 / ((double)k * (double)k), SsN_Tol1);                     // :119
        plan_apd_begin_pot(c, cost, uk, vk, p, q, m, n, phi, b, lk, ak, bk, bk1, wk, huk, wlk);       // :121-122
        SSN_CUDA(cudaMemcpyAsync(lk_new.p, lk.p, sizeof(double) * N1, cudaMemcpyDeviceToDevice, c->stream));
        plan_prox_residual_pot(c, wk, lk_new, p, q, m, n, tk, phi, hp, nullptr, s, tt, scal.p + 4);   // :124-130 (+ s, t for :140)
        SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N1, lk_new.p, hp.p, wlk.p, bk1, Fk.p, mFk.p);
        SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N1, Fk.p, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr,
                   (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, scal.p);
        read_back(c, scal.p, h, 6);                                                                   // h[0] = |Fk|^2, h[4..5] = norm2, count
        double nF = std::sqrt(h[0]), n2_old = h[4], Ecount = h[5];
        double Fk_res = nF;
        int ssn_it = 0;
        while (nF > SsN_Tol) {                                                                       // :136
            ++ssn_it;
            std::swap(lk_old.p, lk_new.p);                                                            // lk_old = lk_new; zk, s, t at lk_old are evaluated already
            Clock ta;
            Csr H0 = asat(c, s, p, q, m, n);                                                          // :146
            if (c->prof) SSN_CUDA(cudaStreamSynchronize(c->stream));
            asat_s += ta.s();
            Clock ts;
            ssn_csr Hv = H0.view();
            ssn_prob_data pd{}; pd.bk1 = bk1; pd.tk = tk; pd.m = m; pd.n = n; pd.p_dev = p; pd.q_dev = q; pd.t_dev = tt.p;
            pd.H0 = &Hv; pd.z_dev = mFk.p; pd.s_dev = s.p; pd.phi_dev = phi;                          // :163-166
            int itl = 0, info[2] = {0, 0}; double resl = 0.0;
            if (inner_solver == 3) pcg4pot(c, &pd, &pcg, zeta.p, &itl, &resl, info);                  // :168
            else { amg4pot(c, &pd, &amg, zeta.p, &itl, &resl, info, inner_solver == 5); ++amg_calls; }   // :171 / :182
            solve_s += ts.s();
            Clock tp;
            SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N1, lk_old.p, (const double*)nullptr, wlk.p, lk_old.p, Fk.p, zeta.p, Fk.p,
                       (const double*)nullptr, scal.p);                                               // :196-197, :203
            read_back(c, scal.p, h, 4);
            const double cFk_old = bk1 / 2 * h[0] - h[1] + 0.5 * tk * n2_old;
            const double ress = std::fabs(h[2]);
            const double nFo = std::sqrt(h[3]);
            const double E_step = Ecount;
            const int ll = pot_linesearch(c, wk, lk_old, zeta, wlk, p, q, m, n, tk, bk1, phi, nu, delta, ll_max, cFk_old, ress, lk_new, hp, s, tt,
                                          scal, &n2_old, &Ecount);                                    // :199-213 (+ the next s, t)
            ls_trials += ll + 1;
            SSN_LAUNCH(c, fk_kernel, gN, 256, 0, N1, lk_new.p, hp.p, wlk.p, bk1, Fk.p, mFk.p);        // :217
            SSN_LAUNCH(c, dots4_kernel, 1, kVT, 0, N1, Fk.p, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr,
                       (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, scal.p);
            read_back(c, scal.p, h, 1);
            nF = std::sqrt(h[0]);
            plan_s += tp.s();
            if (steps_host && nsteps < steps_cap) {
                double* st = steps_host + 7 * nsteps;
                st[0] = k; st[1] = ssn_it; st[2] = E_step; st[3] = info[0]; st[4] = itl; st[5] = ll; st[6] = nF;
            }
            ++nsteps;
            if (op && op->verbose) fprintf(stderr, "   SsN: it=%3d |Fk|=%.2e ll=%3d info=[%d, %d] its=%d res=%.2e E=%.0f\n", ssn_it, nF, ll, info[0], info[1], itl, resl, E_step);
            if (nF <= SsN_Tol) break;                                                                 // :218
            if (std::fabs(nFo - nF) < SsN_Tol) break;                                                 // :224
            if (ssn_it == SsN_IT) break;
            if (Fk_res / nF >= 2) Fk_res = nF;
        }
        ssn_total += ssn_it;
        // :231-238 in one pass: uk1 = prox(zk), vk1, H*uk1, c'xk1 and the four KKT residuals of (uk1, lk_new)
        plan_apd_end_pot(c, cost, wk, uk, lk_new, p, q, m, n, phi, b, tk, ak, uk1, vk1, huk, scal.p + 8);
        double h5[5]; read_back(c, scal.p + 8, h5, 5);
        double cx = h5[0];
        double K[4] = {std::sqrt(h5[1]), std::sqrt(h5[2]), std::sqrt(h5[3]), std::sqrt(h5[4])};
        auto rel = [&](const double (&kk)[4]) {
            double r0 = 0.0;
            for (int i = 0; i < 4; ++i) r0 = std::max(r0, kk[i] / (1 + KK[0][i]));
            return r0;
        };
        if (bk1 < 1e-8 && rel(K) > resk) {                                                           // :253-257: restart
            bk1 = 10 * bk1;                                                                           // uk1 = uk; lk1 = lk; vk1 = uk
            SSN_CUDA(cudaMemcpyAsync(vk.p, uk.p, sizeof(double) * L, cudaMemcpyDeviceToDevice, c->stream));
            kkt(uk, lk, K, cx);
        } else {
            std::swap(uk.p, uk1.p); std::swap(vk.p, vk1.p); std::swap(lk.p, lk_new.p);               // :259
        }
        bk = bk1;
        fx.push_back(cx); KK.push_back({K[0], K[1], K[2], K[3]});
        if (ssn_its_hist) ssn_its_hist[k - 1] = ssn_it;
        rrmax = rel(K);
        if (op && op->verbose) fprintf(stderr, "APD: it=%3d KKT(x,y,z,l)=%.2e %.2e %.2e %.2e fk=%.8e t=%.2fs\n", k, K[0] / (1 + KK[0][0]),
                                       K[1] / (1 + KK[0][1]), K[2] / (1 + KK[0][2]), K[3] / (1 + KK[0][3]), cx, t_loop.s());
        if (rrmax <= KKT_Tol) { converged = 1; break; }                                               // :274
        if (max_outer > 0 && k >= max_outer) break;
        if (max_seconds > 0 && t_loop.s() > max_seconds) break;
    }
    if (k > maxit) k = maxit;
    SSN_CUDA(cudaMemcpyAsync(uk_out, uk.p, sizeof(double) * L, cudaMemcpyDeviceToDevice, c->stream));
    SSN_CUDA(cudaMemcpyAsync(lk_out, lk.p, sizeof(double) * N1, cudaMemcpyDeviceToDevice, c->stream));
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    if (res) {
        res->outer_its = k; res->converged = converged; res->rel_kkt = rrmax; res->objective = fx.back();
        res->ssn_steps = ssn_total; res->ls_trials = ls_trials; res->ls_passes = ls_trials; res->amg_calls = amg_calls;
        res->warmup_s = t_warm; res->loop_s = t_loop.s(); res->solve_s = solve_s; res->asat_s = asat_s; res->plan_s = plan_s;
        res->hist_len = (int)fx.size(); res->steps_len = nsteps;
    }
    for (size_t i = 0; i < fx.size(); ++i) {
        if (fxk_hist) fxk_hist[i] = fx[i];
        if (kkt4_hist) for (int j = 0; j < 4; ++j) kkt4_hist[4 * i + j] = KK[i][j];
    }
}

}  // namespace ssn
