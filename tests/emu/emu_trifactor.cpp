// tests/emu/emu_trifactor.cpp -- runs the REAL csrc/trifactor.cu on the host through tests/emu/common.cuh
// (the test copies trifactor.cu, sparse.cu, amg.cuh and sparse.cuh next to this file before compiling).  Test infrastructure only.
#include "common.cuh"
#include "emu_probe.h"
#include "amg.cuh"
#include "trifactor.cu"          // scans, sorts, transpose: the real sparse.cu, compiled as its own translation unit

namespace {
int fail(const ssn::Error& e, char* err, int errlen) {
    if (err && errlen > 0) { std::strncpy(err, e.msg.c_str(), (size_t)errlen - 1); err[errlen - 1] = 0; }
    return e.code;
}
template <class T> void put(T* dst, const ssn::Buf<T>& b, size_t count) { if (dst && count) std::memcpy(dst, b.p, count * sizeof(T)); }
}

// H: n x n CSR with sorted columns.  Output arrays are sized by the caller for the upper bounds: nnz + n entries,
// n + 1 pointers / level pointers, n rows.  sizes = {nnz(Lf), nnz(Uf), levels of Lf, levels of Uf, has_mid, launches}.
extern "C" int emu_tri_factors(int n, int64_t nnz, const int* ptr, const int* idx, const double* val, int precd, int64_t* sizes,
                               int* lp, int* li, double* lv, int* up, int* ui, double* uv, double* mid,
                               int* lrows, int* llev, int* urows, int* ulev, char* err, int errlen) {
    try {
        ssn_ctx ctx;
        ssn::CsrView H; H.nrows = n; H.ncols = n; H.nnz = nnz; H.ptr = ptr; H.idx = idx; H.val = val;
        ssn::TriFactors F;
        emu::threaded = true;                                // the scans of sparse.cu use __syncthreads and shuffles
        ssn::build_tri_factors_device(&ctx, H, precd, F);
        const size_t nl_ = (size_t)F.lp.p[n], nu_ = (size_t)F.up.p[n];
        sizes[0] = (int64_t)nl_; sizes[1] = (int64_t)nu_; sizes[2] = F.nl; sizes[3] = F.nu; sizes[4] = F.has_mid ? 1 : 0; sizes[5] = ctx.launches;
        put(lp, F.lp, (size_t)n + 1); put(li, F.li, nl_); put(lv, F.lv, nl_);
        put(up, F.up, (size_t)n + 1); put(ui, F.ui, nu_); put(uv, F.uv, nu_);
        if (F.has_mid) put(mid, F.mid, (size_t)n);
        put(lrows, F.lrows, (size_t)n); put(llev, F.llev, (size_t)F.nl + 1);
        put(urows, F.urows, (size_t)n); put(ulev, F.ulev, (size_t)F.nu + 1);
        return 0;
    } catch (const ssn::Error& e) { return fail(e, err, errlen); }
}

// Jk = bk1*I + (T + H0)/tk; H0: N x N CSR (N = m + n); outputs sized N + 1 and nnz + N.  Returns nnz(Jk) in *onnz.
extern "C" int emu_jk_system(int64_t m, int64_t n, int64_t nnz, const int* ptr, const int* idx, const double* val, const double* t,
                             double bk1, double tk, int64_t* onnz, int* optr, int* oidx, double* oval, char* err, int errlen) {
    try {
        ssn_ctx ctx;
        ssn_csr H0; H0.nrows = H0.ncols = m + n; H0.nnz = nnz;
        H0.rowptr_dev = (int32_t*)ptr; H0.colidx_dev = (int32_t*)idx; H0.val_dev = (double*)val;
        ssn_prob_data pd; std::memset(&pd, 0, sizeof(pd));
        pd.bk1 = bk1; pd.tk = tk; pd.m = m; pd.n = n; pd.t_dev = t; pd.H0 = &H0;
        ssn::Csr Jk;
        emu::threaded = true;                                // jk_count_kernel / jk_fill_kernel vote inside the warp
        ssn::jk_system(&ctx, &pd, Jk);
        *onnz = Jk.nnz;
        std::memcpy(optr, Jk.ptr.p, sizeof(int) * (size_t)(m + n + 1));
        std::memcpy(oidx, Jk.idx.p, sizeof(int) * (size_t)Jk.nnz);
        std::memcpy(oval, Jk.val.p, sizeof(double) * (size_t)Jk.nnz);
        return 0;
    } catch (const ssn::Error& e) { return fail(e, err, errlen); }
}
