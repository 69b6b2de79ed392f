// tests/emu/sparse_host.h -- host definitions of the sparse.cu functions trifactor.cu calls (declared by the real
// sparse.cuh / common.cuh): same contracts, plain loops.  For harnesses that do NOT compile the real sparse.cu.
// Test infrastructure only.
#pragma once
#include "sparse.cuh"

namespace ssn {

int64_t scan_counts_to_ptr(ssn_ctx*, const int* counts, int* ptr, int64_t n) {
    int64_t run = 0;
    for (int64_t i = 0; i < n; ++i) { const int v = counts[i]; ptr[i] = (int)run; run += v; }
    ptr[n] = (int)run;
    return run;
}
void exclusive_scan_int(ssn_ctx*, const int* in, int* out, int64_t n) { int run = 0; for (int64_t i = 0; i < n; ++i) { const int v = in[i]; out[i] = run; run += v; } }
void stable_sort_pairs(ssn_ctx*, const int* keys_in, int* keys_out, const int* vals_in, int* vals_out, int64_t n, int key_limit) {
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return keys_in[a] < keys_in[b]; });
    for (int64_t i = 0; i < n; ++i) {
        if (keys_in[order[i]] < 0 || keys_in[order[i]] >= key_limit) throw Error(SSN_E_INVALID, "stable_sort_pairs: key out of range");
        keys_out[i] = keys_in[order[i]]; vals_out[i] = vals_in[order[i]];
    }
}

void iota_int(ssn_ctx*, int* p, int64_t n) { for (int64_t i = 0; i < n; ++i) p[i] = (int)i; }

// At = A' with sorted columns (stable by column: rows ascending inside a column)
Csr transpose(ssn_ctx* c, const CsrView& A) {
    Csr T; T.c = c; T.nrows = A.ncols; T.ncols = A.nrows; T.nnz = A.nnz;
    T.ptr.alloc(c, (size_t)A.ncols + 1); T.idx.alloc(c, A.nnz); T.val.alloc(c, A.nnz);
    std::vector<int> cnt((size_t)A.ncols + 1, 0);
    for (int64_t e = 0; e < A.nnz; ++e) ++cnt[(size_t)A.idx[e] + 1];
    for (int j = 0; j < A.ncols; ++j) cnt[(size_t)j + 1] += cnt[j];
    for (int j = 0; j <= A.ncols; ++j) T.ptr.p[j] = cnt[j];
    for (int i = 0; i < A.nrows; ++i)
        for (int e = A.ptr[i]; e < A.ptr[i + 1]; ++e) { const int j = A.idx[e]; T.idx.p[cnt[j]] = i; T.val.p[cnt[j]] = A.val[e]; ++cnt[j]; }
    return T;
}

Csr csr_alloc_from_counts(ssn_ctx* c, int nrows, int ncols, const int* counts) {
    Csr C; C.c = c; C.nrows = nrows; C.ncols = ncols;
    C.ptr.alloc(c, (size_t)nrows + 1);
    C.nnz = scan_counts_to_ptr(c, counts, C.ptr.p, nrows);
    C.idx.alloc(c, C.nnz); C.val.alloc(c, C.nnz);
    return C;
}

}  // namespace ssn
