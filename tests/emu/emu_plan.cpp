// tests/emu/emu_plan.cpp -- runs the REAL csrc/plan_ops.cu (Ax, Aty, the fused SsN residual, active-set compaction, the
// batched and screened line-search kernels, the Armijo loop, the fused outer-loop updates and warm start) with
// csrc/sparse.cu on the host through tests/emu/common.cuh.  Test infrastructure only.
#include "common.cuh"
#include "emu_probe.h"
#include "sparse.cuh"
#include "plan_ops.cuh"

namespace {
ssn_ctx* g_ctx = nullptr;
std::string g_err;
ssn_ctx* ctx() { if (!g_ctx) g_ctx = new ssn_ctx(); return g_ctx; }
template <class F> int guarded(F f) {
    try { emu::threaded = true; f(); return 0; }
    catch (const ssn::Error& e) { g_err = e.msg; return e.code; }
}
}

extern "C" {
const char* emu_error() { return g_err.c_str(); }
int64_t emu_launches() { return ctx()->launches; }
void emu_set_ls(int max_nt, int screen) { ctx()->ls_max_nt = max_nt; ctx()->ls_screen = screen != 0; }
int emu_ax(const double* x, const double* p, const double* q, int64_t m, int64_t n, double* y) {
    return guarded([&] { ssn::plan_ax(ctx(), x, p, q, m, n, y); });
}
int emu_aty(const double* y, const double* p, const double* q, int64_t m, int64_t n, double* z) {
    return guarded([&] { ssn::plan_aty(ctx(), y, p, q, m, n, z); });
}
int emu_prox_residual(const double* w, const double* lam, const double* p, const double* q, int64_t m, int64_t n, double tk,
                      const double* gama, double gama_s, double* axp, double* prox, double* z, uint8_t* s, double* scal2) {
    return guarded([&] { ssn::plan_prox_residual(ctx(), w, lam, p, q, m, n, tk, gama, gama_s, axp, prox, z, s, scal2); });
}
int emu_prox_residual_pot(const double* w, const double* lam, const double* p, const double* q, int64_t m, int64_t n, double tk,
                          const double* phi, double* hp, double* prox, uint8_t* s, double* t, double* scal3) {
    return guarded([&] { ssn::plan_prox_residual_pot(ctx(), w, lam, p, q, m, n, tk, phi, hp, prox, s, t, scal3); });
}
int emu_active_set(const uint8_t* s, int64_t m, int64_t n, int* colptr, int* yrow, int* ycol, int* rowcount, int64_t* E_out) {
    return guarded([&] {
        ssn::Buf<int> cp, yr, yc, rc;
        const int64_t E = ssn::plan_active_set(ctx(), s, m, n, cp, yr, yc, rc);
        std::memcpy(colptr, cp.p, sizeof(int) * (size_t)(n + 1)); std::memcpy(rowcount, rc.p, sizeof(int) * (size_t)m);
        if (E) { std::memcpy(yrow, yr.p, sizeof(int) * (size_t)E); std::memcpy(ycol, yc.p, sizeof(int) * (size_t)E); }
        *E_out = E;
    });
}
int emu_prox_trials(const double* w, const double* lamT, int nt, const double* p, const double* q, int64_t m, int64_t n, double tk,
                    const double* gama, double gama_s, double* n2_out) {
    return guarded([&] { ssn::plan_prox_trials(ctx(), w, lamT, nt, p, q, m, n, tk, gama, gama_s, n2_out); });
}
int emu_prox_trials_lin(const double* w, const double* lam, const double* zeta, const double* p, const double* q, int64_t m, int64_t n,
                        double tk, double delta, int ll0, int nt, double* out) {
    return guarded([&] { ssn::plan_prox_trials_lin(ctx(), w, lam, zeta, p, q, m, n, tk, delta, ll0, nt, out, nullptr); });
}
int emu_trial_vectors(const double* lam, const double* zeta, const double* wlk, int64_t N, double delta, int ll0, int nt, double* lamT, double* f0) {
    return guarded([&] { ssn::plan_trial_vectors(ctx(), lam, zeta, wlk, N, delta, ll0, nt, lamT, f0); });
}
int emu_linesearch(const double* w, const double* lam_old, const double* zeta, const double* wlk, const double* p, const double* q,
                   int64_t m, int64_t n, double tk, double bk1, const double* gama, double gama_s, double nu, double delta, int ll_max,
                   double cF_old, double ress, int batch, double* lam_new, int* ll_out, double* n2_out, double* cF_out, int* passes_out) {
    return guarded([&] {
        ssn::plan_linesearch(ctx(), w, lam_old, zeta, wlk, p, q, m, n, tk, bk1, gama, gama_s, nu, delta, ll_max, cF_old, ress, batch,
                             lam_new, ll_out, n2_out, cF_out, passes_out);
    });
}
int emu_apd_begin(const double* cost, const double* xk, const double* vk, const double* p, const double* q, int64_t m, int64_t n,
                  double ak, double bk, double* wk, double* axk) {
    return guarded([&] { ssn::plan_apd_begin(ctx(), cost, xk, vk, p, q, m, n, ak, bk, wk, axk); });
}
int emu_apd_end(const double* cost, const double* wk, const double* xk, const double* lam, const double* p, const double* q, int64_t m,
                int64_t n, double tk, double ak, const double* gama, double gama_s, double* xk1, double* vk1, double* axk1, double* scal2) {
    return guarded([&] { ssn::plan_apd_end(ctx(), cost, wk, xk, lam, p, q, m, n, tk, ak, gama, gama_s, xk1, vk1, axk1, scal2); });
}
int emu_warmup_class1(const double* cost, const double* b, const double* p, const double* q, int64_t m, int64_t n, const double* gama,
                      double gama_s, int maxit, double* xk, double* lk) {
    return guarded([&] { ssn::plan_warmup_class1(ctx(), cost, b, p, q, m, n, gama, gama_s, maxit, xk, lk); });
}
int emu_warmup_class2(const double* cost, const double* b, const double* p, const double* q, int64_t m, int64_t n, const double* phi,
                      int maxit, double* uk, double* lk) {
    return guarded([&] { ssn::plan_warmup_class2(ctx(), cost, b, p, q, m, n, phi, maxit, uk, lk); });
}
int emu_apd_begin_pot(const double* cost, const double* uk, const double* vk, const double* p, const double* q, int64_t m, int64_t n,
                      const double* phi, const double* b, const double* lk, double ak, double bk, double bk1, double* wk, double* huk,
                      double* wlk) {
    return guarded([&] { ssn::plan_apd_begin_pot(ctx(), cost, uk, vk, p, q, m, n, phi, b, lk, ak, bk, bk1, wk, huk, wlk); });
}
int emu_apd_end_pot(const double* cost, const double* wk, const double* uk, const double* lk, const double* p, const double* q, int64_t m,
                    int64_t n, const double* phi, const double* b, double tk, double ak, double* uk1, double* vk1, double* huk1,
                    double* scal5) {
    return guarded([&] { ssn::plan_apd_end_pot(ctx(), cost, wk, uk, lk, p, q, m, n, phi, b, tk, ak, uk1, vk1, huk1, scal5); });
}
}
