// tests/emu/emu_solvers.cpp -- runs the REAL csrc/solvers.cu (ASAt assembly, rescaled system, components, closed-form
// inverse) with csrc/sparse.cu and csrc/amg_setup.cu on the host through tests/emu/common.cuh.  What lives in
// translation units that are not emulated is stubbed: the iterative solvers (cooperative kernels, amg_solve.cu) throw,
// the active-set compaction of plan_ops.cu is restated on the host with the same contract.  Test infrastructure only.
#include "common.cuh"
#include "emu_probe.h"
#include "amg.cuh"
#include "solvers.cuh"
#include "plan_ops.cuh"

namespace ssn {
void build_cluster_plan(ssn_ctx*, Hierarchy&) { throw Error(SSN_E_UNSUPPORTED, "emu: build_cluster_plan"); }
void class_amg(ssn_ctx*, const CsrView&, const double*, const AmgOptions&, bool, double*, int*, double*, double*, double*, int*) { throw Error(SSN_E_UNSUPPORTED, "emu: class_amg"); }
void twogrid_bigph(ssn_ctx*, const CsrView&, const double*, const AmgOptions&, double*, int*, double*, double*, double*, int*, bool) { throw Error(SSN_E_UNSUPPORTED, "emu: twogrid_bigph"); }
void pcg_solve(ssn_ctx*, const CsrView&, const double*, const ssn_pcg_options*, double*, int*, double*, double*) { throw Error(SSN_E_UNSUPPORTED, "emu: pcg_solve"); }
void plan_ax(ssn_ctx*, const double*, const double*, const double*, int64_t, int64_t, double*) { throw Error(SSN_E_UNSUPPORTED, "emu: plan_ax"); }
// Y = sparse(reshape(s,m,n)): CSC colptr / yrow (rows ascending inside a column) / ycol, row counts (plan_ops.cu)
int64_t plan_active_set(ssn_ctx* c, const uint8_t* s, int64_t m, int64_t n, Buf<int>& colptr, Buf<int>& yrow, Buf<int>& ycol, Buf<int>& rowcount) {
    int64_t E = 0;
    for (int64_t k = 0; k < m * n; ++k) E += s[k] ? 1 : 0;
    colptr.alloc(c, (size_t)n + 1); yrow.alloc(c, (size_t)E); ycol.alloc(c, (size_t)E); rowcount.alloc(c, (size_t)m); rowcount.zero();
    int64_t e = 0;
    for (int64_t j = 0; j < n; ++j) {
        colptr.p[j] = (int)e;
        for (int64_t i = 0; i < m; ++i) if (s[j * m + i]) { yrow.p[e] = (int)i; ycol.p[e] = (int)j; ++rowcount.p[i]; ++e; }
    }
    colptr.p[n] = (int)e;
    return E;
}
}

namespace {
ssn_ctx* g_ctx = nullptr;
ssn::Csr g_out;
std::string g_err;
ssn_ctx* ctx() { if (!g_ctx) { g_ctx = new ssn_ctx(); } return g_ctx; }
template <class F> int guarded(F f) {
    try { emu::threaded = true; f(); return 0; }
    catch (const ssn::Error& e) { g_err = e.msg; return e.code; }
}
}

extern "C" {
const char* emu_error() { return g_err.c_str(); }
int64_t emu_host_reads() { return emu::host_reads; }
void emu_sizes(int, int64_t* out) { out[0] = g_out.nrows; out[1] = g_out.ncols; out[2] = g_out.nnz; }
void emu_fetch(int, int* ptr, int* idx, double* val) {
    std::memcpy(ptr, g_out.ptr.p, sizeof(int) * (size_t)(g_out.nrows + 1));
    std::memcpy(idx, g_out.idx.p, sizeof(int) * (size_t)g_out.nnz);
    std::memcpy(val, g_out.val.p, sizeof(double) * (size_t)g_out.nnz);
}
int emu_asat(const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n) {
    return guarded([&] { g_out = ssn::asat(ctx(), s, p, q, m, n); });
}
int emu_asat_coo(const long long* lin, int64_t E, const double* p, const double* q, int64_t m, int64_t n) {
    return guarded([&] { g_out = ssn::asat_coo(ctx(), lin, E, p, q, m, n); });
}
int emu_asatz(const double* z, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n, double* y) {
    return guarded([&] { ssn::asatz(ctx(), z, s, p, q, m, n, y); });
}
int emu_components(int64_t n, int64_t ncols, int64_t nnz, const int* ap, const int* ai, const double* av, int* blocks, int* sizes, int* perm, int* r, int* ncomp) {
    return guarded([&] {
        ssn::CsrView A; A.nrows = (int)n; A.ncols = (int)ncols; A.nnz = nnz; A.ptr = ap; A.idx = ai; A.val = av;
        ssn::components(ctx(), A, blocks, sizes, perm, r, ncomp);
    });
}
// Ae -> g_out, f -> f_out
int emu_rescaled_system(double bk1, double tk, int64_t m, int64_t n, const double* p, const double* q, const double* t, const double* z,
                        int64_t nnz, const int* hp, const int* hi, const double* hv, double* f_out) {
    return guarded([&] {
        ssn_csr H0; H0.nrows = H0.ncols = m + n; H0.nnz = nnz; H0.rowptr_dev = (int32_t*)hp; H0.colidx_dev = (int32_t*)hi; H0.val_dev = (double*)hv;
        ssn_prob_data pd; std::memset(&pd, 0, sizeof(pd));
        pd.bk1 = bk1; pd.tk = tk; pd.m = m; pd.n = n; pd.p_dev = p; pd.q_dev = q; pd.t_dev = t; pd.H0 = &H0; pd.z_dev = z;
        ssn::Buf<double> qp, Kd;
        ssn::rescaled_system(ctx(), &pd, g_out, f_out, qp, Kd);
    });
}
int emu_invaat(const double* x, const double* p, const double* q, int64_t m, int64_t n, double sg1, double sg2, double* y) {
    return guarded([&] { ssn::invaat(ctx(), x, p, q, m, n, sg1, sg2, y); });
}
}
