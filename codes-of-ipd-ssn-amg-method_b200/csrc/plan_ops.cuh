// plan_ops.cuh -- internal interface of plan_ops.cu
#pragma once
#include "common.cuh"
namespace ssn {
void plan_ax(ssn_ctx* c, const double* x, const double* p, const double* q, int64_t m, int64_t n, double* y);
void plan_aty(ssn_ctx* c, const double* y, const double* p, const double* q, int64_t m, int64_t n, double* z);
// scal2_dev receives {||prox(z)||^2, nnz(s)} (device, 2 doubles)
void plan_prox_residual(ssn_ctx* c, const double* w, const double* lam, const double* p, const double* q,
                        int64_t m, int64_t n, double tk, const double* gama, double gama_s, double* axp_out,
                        double* prox_out, double* z_out, uint8_t* s_out, double* scal2_dev);
// batched line-search trials and the Armijo loop built on them (plan_ops.cu)
void plan_prox_trials(ssn_ctx* c, const double* w, const double* lamT, int nt, const double* p, const double* q,
                      int64_t m, int64_t n, double tk, const double* gama, double gama_s, double* n2_out_dev,
                      const int* nonunit_dev = nullptr);
void plan_prox_trials_lin(ssn_ctx* c, const double* w, const double* lam, const double* zeta, const double* p, const double* q,
                          int64_t m, int64_t n, double tk, double delta, int ll0, int nt, double* out_dev, const int* nonunit_dev);
void plan_trial_vectors(ssn_ctx* c, const double* lam, const double* zeta, const double* wlk, int64_t N, double delta,
                        int ll0, int nt, double* lamT, double* f0_out);
void plan_linesearch(ssn_ctx* c, const double* w, const double* lam_old, const double* zeta, const double* wlk,
                     const double* p, const double* q, int64_t m, int64_t n, double tk, double bk1, const double* gama,
                     double gama_s, double nu, double delta, int ll_max, double cF_old, double ress, int batch,
                     double* lam_new, int* ll_out, double* n2_out, double* cF_out, int* passes_out);
void plan_warmup_class1(ssn_ctx* c, const double* cost, const double* b, const double* p, const double* q, int64_t m,
                        int64_t n, const double* gama, double gama_s, int maxit, double* xk_out, double* lk_out);
void plan_warm_stage(ssn_ctx* c, int stage, double* xk, double* vk, double* wk, double* pik, double* lk2, double* dd,
                     const double* cost, const double* p, const double* q, const double* b, const double* lk1, const double* axk,
                     const double* y, int64_t m, int64_t n, const double* gama, double gama_s, double ak, double bk, double gk,
                     double* out1, double* out2);
void plan_apd_begin(ssn_ctx* c, const double* cost, const double* xk, const double* vk, const double* p, const double* q,
                    int64_t m, int64_t n, double ak, double bk, double* wk_out, double* axk_out);
void plan_apd_end(ssn_ctx* c, const double* cost, const double* wk, const double* xk, const double* lam, const double* p,
                  const double* q, int64_t m, int64_t n, double tk, double ak, const double* gama, double gama_s, double* xk1,
                  double* vk1, double* axk1_out, double* scal2_dev);
void plan_prox_residual_pot(ssn_ctx* c, const double* w, const double* lam, const double* p, const double* q, int64_t m, int64_t n,
                            double tk, const double* phi, double* hp_out, double* prox_out, uint8_t* s_out, double* t_out,
                            double* scal3_dev);
// Class 2 (partial OT), u = [x (m*n) ; y (n) ; z (m)], b = [r ; l ; mu], lk of n+m+1 entries
void plan_warmup_class2(ssn_ctx* c, const double* cost, const double* b, const double* p, const double* q, int64_t m, int64_t n,
                        const double* phi, int maxit, double* uk_out, double* lk_out);
void plan_apd_begin_pot(ssn_ctx* c, const double* cost, const double* uk, const double* vk, const double* p, const double* q, int64_t m,
                        int64_t n, const double* phi, const double* b, const double* lk, double ak, double bk, double bk1, double* wk_out,
                        double* huk_out, double* wlk_out);
void plan_apd_end_pot(ssn_ctx* c, const double* cost, const double* wk, const double* uk, const double* lk, const double* p, const double* q,
                      int64_t m, int64_t n, const double* phi, const double* b, double tk, double ak, double* uk1, double* vk1,
                      double* huk1_out, double* scal5_dev);
int64_t plan_active_set(ssn_ctx* c, const uint8_t* s, int64_t m, int64_t n, Buf<int>& colptr, Buf<int>& yrow,
                        Buf<int>& ycol, Buf<int>& rowcount);
}  // namespace ssn
