// sparse.cuh -- internal interface of sparse.cu (CSR building blocks of the AMG path)
#pragma once
#include "common.cuh"

namespace ssn {

// y = A*x (solve-phase kernel: summation order free)
void spmv(ssn_ctx* c, const CsrView& A, const double* x, double* y);
// y += A*x
void spmv_add(ssn_ctx* c, const CsrView& A, const double* x, double* y);
// At = A' with sorted columns (stable radix sort by column: deterministic)
Csr transpose(ssn_ctx* c, const CsrView& A);
// C = A*B in the frozen order (row-wise Gustavson, k ascending, multiply-then-add, no FMA;
// bit-identical to the oracle's column-wise Gustavson on the transposed problem); exact
// zeros are dropped from the result like a MATLAB sparse product.
Csr spgemm(ssn_ctx* c, const CsrView& A, const CsrView& B);
// C = A + alpha*B (union pattern, exact zeros dropped)
Csr sparse_add(ssn_ctx* c, const CsrView& A, double alpha, const CsrView& B);
// rows/cols gather: C = A(sel,sel) where newidx[v] = position of node v in sel or -1;
// sel ascending (so column order is preserved)
Csr extract_principal(ssn_ctx* c, const CsrView& A, const int* sel, int nsel, const int* newidx);
// drop entries with val == 0
Csr drop_zeros(ssn_ctx* c, const CsrView& A);
// diag[i] = A(i,i) (0 if absent)
void extract_diag(ssn_ctx* c, const CsrView& A, double* diag);
// rowidx[e] = row of entry e
void expand_rows(ssn_ctx* c, const CsrView& A, int* rowidx);
// copy
Csr csr_copy(ssn_ctx* c, const CsrView& A);
// allocate a CSR from per-row counts (fills ptr, allocates idx/val)
Csr csr_alloc_from_counts(ssn_ctx* c, int nrows, int ncols, const int* counts);
// the same when the caller knows the total already (no host read)
Csr csr_alloc_known(ssn_ctx* c, int nrows, int ncols, const int* counts, int64_t nnz);

// generic device helpers
void fill_double(ssn_ctx* c, double* p, int64_t n, double v);
void fill_int(ssn_ctx* c, int* p, int64_t n, int v);
void iota_int(ssn_ctx* c, int* p, int64_t n);
// deterministic sum / dot / norm (two-stage, fixed order); result on host (synchronises)
double dev_sum(ssn_ctx* c, const double* x, int64_t n);
// the same sum (same order, same value) left in *out_dev: no host read
void dev_sum_async(ssn_ctx* c, const double* x, int64_t n, double* out_dev);
double dev_dot(ssn_ctx* c, const double* x, const double* y, int64_t n);
int64_t dev_count_nonzero_u8(ssn_ctx* c, const uint8_t* x, int64_t n);

}  // namespace ssn
