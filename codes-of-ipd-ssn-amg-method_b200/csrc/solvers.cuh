// solvers.cuh -- internal interface of solvers.cu
#pragma once
#include "common.cuh"
#include "sparse.cuh"
namespace ssn {
Csr asat(ssn_ctx* c, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n);
Csr asat_coo(ssn_ctx* c, const long long* lin_sorted, int64_t E, const double* p, const double* q, int64_t m, int64_t n);
void active_coo(ssn_ctx* c, const uint8_t* s, int64_t m_loc, int64_t n, int64_t row_offset, int64_t m_global, long long** lin_out, int64_t* E_out);
void asatz(ssn_ctx* c, const double* z, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n, double* y);
void invaat(ssn_ctx* c, const double* x, const double* p, const double* q, int64_t m, int64_t n, double sg1, double sg2, double* y);
void invhht(ssn_ctx* c, const double* v, const double* p, const double* q, int64_t m, int64_t n, double sg, const double* phi, double* y);
void amg4pot(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, double* zeta, int* it, double* res, int* info,
             bool twogrid = false);
void pcg4pot(ssn_ctx* c, const ssn_prob_data* pd, const ssn_pcg_options* opts, double* zeta, int* it, double* res, int* info);
// the Class 1 script as one call (apd_driver.cu)
void apd_ssn_class1(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                    int64_t n, const double* gama, double gama_s, const ssn_apd_options* op, double* xk_out, double* lk_out,
                    ssn_apd_result* res, double* fxk_hist, double* kktx_hist, double* kktl_hist, int32_t* ssn_its_hist,
                    double* steps_host, int64_t steps_cap);
void ssn_step_class1(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                     int64_t n, double bk1, double tk, const double* gama, double gama_s, int inner_solver, const ssn_amg_options* amg_in,
                     double* lk_new, double* Fk_new, double* info12);
// Class 2 (partial OT) as one call / one step (apd_driver.cu)
void apd_ssn_class2(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                    int64_t n, double mu, const double* phi, const ssn_apd_options* op, double* uk_out, double* lk_out,
                    ssn_apd_result* res, double* fxk_hist, double* kkt4_hist, int32_t* ssn_its_hist, double* steps_host, int64_t steps_cap);
void ssn_step_class2(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                     int64_t n, double bk1, double tk, const double* phi, int inner_solver, const ssn_amg_options* amg_in,
                     const ssn_pcg_options* pcg_in, double* lk_new, double* Fk_new, double* info12);
}  // namespace ssn
