// tests/emu/emu_amg.cpp -- runs the REAL csrc/sparse.cu and csrc/amg_setup.cu on the host through tests/emu/common.cuh
// (the test copies them, with amg.cuh / sparse.cuh, next to this file; cub is tests/emu/cub).  Test infrastructure only.
#include "common.cuh"
#include "emu_probe.h"
#include "amg.cuh"

namespace ssn {
// referenced by amg_setup.cu, defined in translation units that are not emulated (never reached: no_cluster = true)
void build_cluster_plan(ssn_ctx*, Hierarchy&) { throw Error(SSN_E_UNSUPPORTED, "emu: build_cluster_plan"); }
}

namespace {
ssn_ctx* g_ctx = nullptr;
ssn::Csr g_out[4];                                  // results of the last call, fetched with emu_fetch
std::vector<uint8_t> g_flags[2];
std::string g_err;
int g_bigph = 1;                                    // amg_options.bigph of emu_amg_setup

ssn_ctx* ctx() { if (!g_ctx) { g_ctx = new ssn_ctx(); emu::threaded = true; ssn::rng_reset(g_ctx, 5489u); } return g_ctx; }
ssn::CsrView view(int64_t nr, int64_t nc, int64_t nnz, const int* ptr, const int* idx, const double* val) {
    ssn::CsrView A; A.nrows = (int)nr; A.ncols = (int)nc; A.nnz = nnz; A.ptr = ptr; A.idx = idx; A.val = val; return A;
}
template <class F> int guarded(F f) {
    try { emu::threaded = true; f(); return 0; }
    catch (const ssn::Error& e) { g_err = e.msg; return e.code; }
}
}

extern "C" {
const char* emu_error() { return g_err.c_str(); }
void emu_sizes(int k, int64_t* out) { out[0] = g_out[k].nrows; out[1] = g_out[k].ncols; out[2] = g_out[k].nnz; }
void emu_fetch(int k, int* ptr, int* idx, double* val) {
    std::memcpy(ptr, g_out[k].ptr.p, sizeof(int) * (size_t)(g_out[k].nrows + 1));
    std::memcpy(idx, g_out[k].idx.p, sizeof(int) * (size_t)g_out[k].nnz);
    std::memcpy(val, g_out[k].val.p, sizeof(double) * (size_t)g_out[k].nnz);
}
void emu_fetch_flags(int k, uint8_t* out) { std::memcpy(out, g_flags[k].data(), g_flags[k].size()); }
int64_t emu_launches() { return ctx()->launches; }
int64_t emu_host_reads() { return emu::host_reads; }
// "phase launches host_reads" lines of everything run so far (static buffer)
const char* emu_phase_counts() {
    static std::string out; out.clear();
    for (auto& kv : ssn::phase_counts) out += kv.first + "\t" + std::to_string(kv.second.first) + "\t" + std::to_string(kv.second.second) + "\n";
    return out.c_str();
}
void emu_phase_reset() { ssn::phase_counts.clear(); }
void emu_set_bigph(int v) { g_bigph = v; }
void emu_set_small_scan_max(int v) { ctx()->small_scan_max = v; }

// Class_AMG's setup phase (AMG/Class_AMG.m:41-85): returns the number of levels; level k's matrix, prolongation and
// ones'*A_k*ones are fetched with emu_level
int emu_amg_setup(int64_t n, int64_t nnz, const int* ap, const int* ai, const double* av, double theta, int smoth, int isnsp, int fnode, int* levels) {
    return guarded([&] {
        ssn_amg_options o; std::memset(&o, 0, sizeof(o));
        o.retol = 1e-11; o.maxit = 30; o.smoth = smoth; o.cycle = 'w'; o.theta = theta; o.bigph = g_bigph; o.inter = 1; o.isnsp = isnsp; o.fnode = fnode;
        ssn::amg_setup(ctx(), view(n, n, nnz, ap, ai, av), ssn::resolve_options(&o));
        *levels = ctx()->hier->J;
    });
}
// which = 0: A_k -> g_out[0]; 1: Pro_k -> g_out[0] (k >= 1); returns xx of the level
double emu_level(int k, int which) {
    ssn::Level& L = ctx()->hier->lv[(size_t)k];
    g_out[0] = ssn::csr_copy(ctx(), which == 0 ? ssn::CsrView(L.A) : ssn::CsrView(L.P));
    return L.xx;
}
void emu_amg_clear() { ssn::amg_clear(ctx()); }

// Class_AMG's solve loop on the live hierarchy through amg_cluster.cu (a cluster of 16 emulated CTAs): level kd is applied
// as the dense cycle operator B (N_kd x N_kd, row-major, made by the test from the oracle's cycle).  status: -1 the
// hierarchy did not qualify, else the kernel's it_out[2]
static int g_last_halo = -1;
int emu_last_halo() { return g_last_halo; }
int emu_dsm_solve(int kd, const double* B, const double* b, const double* guess, int isnsp, int wcycle, double retol, int maxit,
                  double* x_out, int* it, double* relk, double* rho, int* hist_len, int* status) {
    return guarded([&] {
        ssn::Hierarchy& H = *ctx()->hier;
        ssn::Level& L = H.lv[(size_t)kd];
        L.B.alloc(ctx(), (size_t)L.N * L.N);
        std::memcpy(L.B.p, B, sizeof(double) * (size_t)L.N * L.N);
        H.dense_from = kd;
        const int n = H.lv[0].N, hl = maxit + 2;
        ssn::Buf<double> x(ctx(), (size_t)n), bb(ctx(), (size_t)n), hist(ctx(), (size_t)2 * hl);
        ssn::Buf<int> iout(ctx(), 4);
        std::memcpy(x.p, guess, sizeof(double) * n); std::memcpy(bb.p, b, sizeof(double) * n);
        ssn::AmgOptions o{}; o.retol = retol; o.maxit = maxit; o.isnsp = isnsp;
        const bool ok = ssn::dsm_cluster_solve(ctx(), H, bb.p, x.p, o, wcycle != 0, hist.p, hl, iout.p);
        *status = ok ? iout.p[2] : -1;
        g_last_halo = ok ? iout.p[3] : -1;
        if (!ok || iout.p[2] != 0) return;
        *it = iout.p[0]; *hist_len = iout.p[1];
        std::memcpy(relk, hist.p, sizeof(double) * iout.p[1]); std::memcpy(rho, hist.p + hl, sizeof(double) * iout.p[1]);
        std::memcpy(x_out, x.p, sizeof(double) * n);
    });
}

// twogrid_bigph (AMG/twogrid_bigph.m) through the cluster kernel: the two-level hierarchy (amg_setup with max_levels = 2) and the
// whole iteration loop with the coarse PCG (retol [] -> 1e-11, maxit 100, Jacobi; :98-99) inside dsm_solve_kernel
int emu_twogrid_dsm(int64_t n, int64_t nnz, const int* ap, const int* ai, const double* av, int smoth, int isnsp, int fnode, const double* b,
                    const double* guess, double retol, int maxit, double* x_out, int* it, double* relk, double* rho, int* hist_len, int* status) {
    return guarded([&] {
        ssn_amg_options oo; std::memset(&oo, 0, sizeof(oo));
        oo.retol = retol; oo.maxit = maxit; oo.smoth = smoth; oo.cycle = 'v'; oo.theta = 0.25; oo.bigph = 1; oo.inter = 1; oo.isnsp = isnsp; oo.fnode = fnode;
        const ssn::AmgOptions o = ssn::resolve_options(&oo);
        ssn::amg_setup(ctx(), view(n, n, nnz, ap, ai, av), o, 2);
        ssn::Hierarchy& H = *ctx()->hier;
        const int N = H.lv[0].N, hl = maxit + 2;
        ssn::Buf<double> x(ctx(), (size_t)N), bb(ctx(), (size_t)N), hist(ctx(), (size_t)2 * hl);
        ssn::Buf<int> iout(ctx(), 4);
        std::memcpy(x.p, guess, sizeof(double) * N); std::memcpy(bb.p, b, sizeof(double) * N);
        ssn_pcg_options po{}; po.retol = -1.0; po.maxit = 100; po.precd = 2; po.nf = 0; po.guess_dev = nullptr;
        const bool ok = ssn::dsm_cluster_solve(ctx(), H, bb.p, x.p, o, false, hist.p, hl, iout.p, &po);
        *status = ok ? iout.p[2] : -1;
        if (ok && iout.p[2] == 0) {
            *it = iout.p[0]; *hist_len = iout.p[1];
            std::memcpy(relk, hist.p, sizeof(double) * iout.p[1]); std::memcpy(rho, hist.p + hl, sizeof(double) * iout.p[1]);
            std::memcpy(x_out, x.p, sizeof(double) * N);
        }
        ssn::amg_clear(ctx());
    });
}

int emu_rng_reset() { return guarded([&] { ssn::rng_reset(ctx(), 5489u); }); }
int64_t emu_rng_drawn() { return ctx()->rng_drawn; }
int emu_rand(int64_t count, double* out) { return guarded([&] { ssn::rng_rand(ctx(), count, out); }); }

void emu_set_spgemm_slab(int64_t limit) { ctx()->spgemm_slab_limit = limit; }
int emu_spgemm(int64_t ar, int64_t ac, int64_t annz, const int* ap, const int* ai, const double* av,
               int64_t br, int64_t bc, int64_t bnnz, const int* bp, const int* bi, const double* bv) {
    return guarded([&] { g_out[0] = ssn::spgemm(ctx(), view(ar, ac, annz, ap, ai, av), view(br, bc, bnnz, bp, bi, bv)); });
}
int emu_transpose(int64_t ar, int64_t ac, int64_t annz, const int* ap, const int* ai, const double* av) {
    return guarded([&] { g_out[0] = ssn::transpose(ctx(), view(ar, ac, annz, ap, ai, av)); });
}
int emu_sparse_add(int64_t ar, int64_t ac, int64_t annz, const int* ap, const int* ai, const double* av, double alpha,
                   int64_t bnnz, const int* bp, const int* bi, const double* bv) {
    return guarded([&] { g_out[0] = ssn::sparse_add(ctx(), view(ar, ac, annz, ap, ai, av), alpha, view(ar, ac, bnnz, bp, bi, bv)); });
}
int emu_strength(int64_t n, int64_t nnz, const int* ap, const int* ai, const double* av, int which) {
    return guarded([&] { g_out[0] = ssn::strength_matrix(ctx(), view(n, n, nnz, ap, ai, av), which == 1 ? 1 : 2); });
}
// [isC, isF, As] = mis_set(A, theta): flags 0 / 1 fetched with emu_fetch_flags, As = g_out[0]
int emu_mis_set(int64_t n, int64_t nnz, const int* ap, const int* ai, const double* av, double theta) {
    return guarded([&] {
        ssn::CsrView A = view(n, n, nnz, ap, ai, av);
        ssn::Buf<uint8_t> isC(ctx(), (size_t)n), isF(ctx(), (size_t)n), flags;
        ssn::mis_set(ctx(), A, theta, isC.p, isF.p, flags);
        g_out[0] = ssn::flags_to_csr(ctx(), A, flags.p);
        g_flags[0].assign(isC.p, isC.p + n); g_flags[1].assign(isF.p, isF.p + n);
    });
}
// [Ac, Pro, As, indC] = transfer(A, amg_options) with global J = level_J: Ac = g_out[0], Pro = g_out[1], As = g_out[2]
int emu_transfer(int64_t n, int64_t nnz, const int* ap, const int* ai, const double* av, double theta, int bigph, int inter,
                 int isnsp, int fnode, int level_J) {
    return guarded([&] {
        ssn::CsrView A = view(n, n, nnz, ap, ai, av);
        ssn_amg_options o; std::memset(&o, 0, sizeof(o));
        o.retol = -1; o.maxit = -1; o.smoth = -1; o.cycle = -1;
        o.theta = theta; o.bigph = bigph; o.inter = inter; o.isnsp = isnsp; o.fnode = fnode;
        ssn::AmgOptions ro = ssn::resolve_options(&o);
        ssn::Csr ac, pro; ssn::Buf<uint8_t> isC, flags;
        ssn::transfer(ctx(), A, ro, level_J, ac, pro, &isC, &flags);
        g_out[2] = ssn::flags_to_csr(ctx(), A, flags.p);
        g_flags[0].assign(isC.p, isC.p + n);
        g_out[0] = std::move(ac); g_out[1] = std::move(pro);
    });
}
}
