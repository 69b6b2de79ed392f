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
int64_t plan_active_set(ssn_ctx* c, const uint8_t* s, int64_t m, int64_t n, Buf<int>& colptr, Buf<int>& yrow,
                        Buf<int>& ycol, Buf<int>& rowcount);
}  // namespace ssn
