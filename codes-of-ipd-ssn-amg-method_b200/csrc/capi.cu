// capi.cu -- the extern "C" boundary of libssnamg.so (include/ssnamg.h).  Every entry point
// converts internal exceptions into status codes; nothing C++ crosses the ABI.
#include "amg.cuh"
#include <mutex>
#include "plan_ops.cuh"
#include "solvers.cuh"
#include <cstdlib>

using namespace ssn;

namespace {

template <class F>
int guarded(ssn_ctx* ctx, F&& f) {
    if (!ctx) return SSN_E_INVALID;
    try {
        SSN_CUDA(cudaSetDevice(ctx->device));
        f();
        return SSN_OK;
    } catch (const Error& e) {
        ctx->err = e.msg;
        cudaGetLastError();
        return e.code;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        return SSN_E_INVALID;
    } catch (...) {
        ctx->err = "unknown failure";
        return SSN_E_INVALID;
    }
}

void sync(ssn_ctx* c) { SSN_CUDA(cudaStreamSynchronize(c->stream)); }

}  // namespace

extern "C" {

int ssn_version(void) { return 100; }

int ssn_create(ssn_ctx** out, int device) {
    if (!out) return SSN_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SSN_E_CUDA;   // no CPU fallback
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return SSN_E_CUDA; }
    if (device >= ndev) return SSN_E_INVALID;
    ssn_ctx* c = new ssn_ctx();
    c->device = device;
    // the 8-CTA cluster cycle kernel is correct (parity-tested) but not yet faster than the single-CTA
    // kernel at these level sizes: opt-in with SSN_CLUSTER=1 until its per-phase latency is tuned
    { const char* e = getenv("SSN_CLUSTER"); c->no_cluster = !(e && e[0] == '1'); }
    { const char* e = getenv("SSN_PERSIST"); c->persist = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_PERSIST_MAXNNZ"); if (e && atoll(e) > 0) c->persist_max_nnz = atoll(e); }
    { const char* e = getenv("SSN_CLUSTER_SOLVE"); c->cluster_solve = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_STAGE_DENSE"); c->stage_dense = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_MIS_CLUSTER"); c->mis_cluster = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_DSM_SOLVE"); c->dsm_solve = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_FUSED_SETUP"); c->fused_setup = (e && e[0] == '1'); }
    { const char* e = getenv("SSN_CLUSTER_MAXNNZ"); if (e && atoll(e) > 0) c->cluster_max_nnz = atoll(e); }
    { const char* e = getenv("SSN_LS_MAXNT"); if (e && atoi(e) >= 8) c->ls_max_nt = atoi(e) > 128 ? 128 : atoi(e); }
    { const char* e = getenv("SSN_LS_SCREEN"); c->ls_screen = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_PLAN_STAGE"); c->plan_stage = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_PLAN_WAVES"); if (e && atoi(e) >= 1 && atoi(e) <= 16) c->plan_waves = atoi(e); }
    { const char* e = getenv("SSN_DENSE_TAIL"); c->dense_tail = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_SMALL_SCAN_MAX"); if (e && atoi(e) >= 0) c->small_scan_max = atoi(e); }
    { const char* e = getenv("SSN_DEVICE_SETUP"); c->device_setup = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_TG_CLUSTER"); c->tg_cluster = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_BUF_CACHE"); c->buf_cache = !(e && e[0] == '0'); }
    { const char* e = getenv("SSN_DENSE_MAXN"); if (e && atoi(e) > 0) c->dense_max_n = atoi(e); }
    try {
        SSN_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        SSN_CUDA(cudaGetDeviceProperties(&prop, device));
        c->num_sms = prop.multiProcessorCount;
        c->smem_optin = prop.sharedMemPerBlockOptin;
        cudaMemPool_t pool;
        SSN_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX;
        SSN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        SSN_CUDA(cudaMallocHost((void**)&c->h_pin, sizeof(double) * ssn_ctx::kPinDoubles));
        SSN_CUDA(cudaMallocHost((void**)&c->h_up, (size_t)ssn_ctx::kUpSlots * ssn_ctx::kUpBytes));
        { const char* e = getenv("SSN_POLL_READS"); c->poll_reads = !(e && e[0] == '0'); }
        if (cudaHostAlloc((void**)&c->h_poll, sizeof(double) * (ssn_ctx::kPinDoubles + 2), cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer((void**)&c->d_poll, c->h_poll, 0) == cudaSuccess) {
            std::memset(c->h_poll, 0, sizeof(double) * (ssn_ctx::kPinDoubles + 2));
        } else { (void)cudaGetLastError(); if (c->h_poll) cudaFreeHost(c->h_poll); c->h_poll = nullptr; c->d_poll = nullptr; }
        SSN_CUDA(cudaMalloc((void**)&c->mt_state, sizeof(uint32_t) * 625));
        rng_reset(c, 5489u);
        SSN_CUDA(cudaStreamSynchronize(c->stream));
    } catch (const Error& e) {
        fprintf(stderr, "ssn_create: %s\n", e.msg.c_str());
        delete c;
        return e.code;
    }
    *out = c;
    return SSN_OK;
}

int ssn_destroy(ssn_ctx* c) {
    if (!c) return SSN_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->hier) { delete c->hier; c->hier = nullptr; }
    ssn::ctx_cache_purge(c);
    cudaStreamSynchronize(c->stream);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_up) cudaFreeHost(c->h_up);
    if (c->h_poll) cudaFreeHost(c->h_poll);
    if (c->mt_state) cudaFree(c->mt_state);
    delete c;
    return SSN_OK;
}

// process-wide default context (the MEX shims' shared hidden state: hierarchy handle + MATLAB random stream)
static std::mutex g_default_mu;
static ssn_ctx* g_default_ctx = nullptr;
static int g_default_refs = 0;

int ssn_default_ctx_acquire(ssn_ctx** out) {
    if (!out) return SSN_E_INVALID;
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_ctx) {
        const int st = ssn_create(&g_default_ctx, -1);
        if (st != SSN_OK) { g_default_ctx = nullptr; *out = nullptr; return st; }
        g_default_refs = 0;
    }
    ++g_default_refs;
    *out = g_default_ctx;
    return SSN_OK;
}
int ssn_default_ctx_release(void) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (g_default_refs <= 0) return SSN_E_INVALID;
    if (--g_default_refs == 0) { ssn_destroy(g_default_ctx); g_default_ctx = nullptr; }
    return SSN_OK;
}
int ssn_default_ctx_refcount(void) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    return g_default_refs;
}

const char* ssn_last_error(ssn_ctx* c) { return c ? c->err.c_str() : "null context"; }
int ssn_set_stream(ssn_ctx* c, void* s) { if (!c) return SSN_E_INVALID; c->stream = (cudaStream_t)s; return SSN_OK; }
int ssn_synchronize(ssn_ctx* c) { return guarded(c, [&] { sync(c); }); }
int64_t ssn_launch_count(ssn_ctx* c) { return c ? c->launches : 0; }

int ssn_kernel_timer(ssn_ctx* c, int on) { if (!c) return SSN_E_INVALID; c->ktimer = on != 0; c->kt_ms = 0.0; c->kt_n = 0; return SSN_OK; }
int ssn_kernel_timer_read(ssn_ctx* c, double* total_ms, int64_t* launches) {
    if (!c) return SSN_E_INVALID;
    if (total_ms) *total_ms = c->kt_ms;
    if (launches) *launches = c->kt_n;
    return SSN_OK;
}
int ssn_debug_barrier_bench(ssn_ctx* c, int iters, int which, double* cycles_per_barrier) {
    return guarded(c, [&] { *cycles_per_barrier = (which % 100 >= 10) ? dsm_bench(c, iters, which) : barrier_bench(c, iters, which); });
}
int ssn_set_persistent(ssn_ctx* c, int on) { if (!c) return SSN_E_INVALID; c->persist = on != 0; return SSN_OK; }
int ssn_set_device_setup(ssn_ctx* c, int on) { if (!c) return SSN_E_INVALID; c->device_setup = on != 0; return SSN_OK; }
int ssn_set_fused_setup(ssn_ctx* c, int on) { if (!c) return SSN_E_INVALID; c->fused_setup = on != 0; return SSN_OK; }
int ssn_set_cluster_solve(ssn_ctx* c, int on) { if (!c) return SSN_E_INVALID; c->cluster_solve = on != 0; c->dsm_solve = on != 1; return SSN_OK; }
int ssn_set_spgemm_slab_limit(ssn_ctx* c, int64_t limit) {
    if (!c || limit < 1 || limit > ((int64_t)1 << 30)) return SSN_E_INVALID;
    c->spgemm_slab_limit = limit;
    return SSN_OK;
}
int ssn_set_dense_tail(ssn_ctx* c, int dense_tail, int dense_max_n) {
    if (!c) return SSN_E_INVALID;
    c->dense_tail = dense_tail != 0;
    if (dense_max_n > 0) c->dense_max_n = dense_max_n;
    return SSN_OK;
}
int ssn_profile_enable(ssn_ctx* c, int on) { if (!c) return SSN_E_INVALID; c->prof = on != 0; return SSN_OK; }
const char* ssn_profile_dump(ssn_ctx* c) {
    if (!c) return "";
    c->prof_text.clear();
    char line[256];
    for (auto& kv : c->prof_acc) {
        snprintf(line, sizeof(line), "%-44s %12.3f  calls %8ld\n", kv.first.c_str(), kv.second.first, kv.second.second);
        c->prof_text += line;
    }
    snprintf(line, sizeof(line), "%-44s %12.0f  misses %8ld\n", "device buffers requested since the last dump", (double)c->buf_allocs, (long)c->buf_misses);
    c->prof_text += line;
    c->buf_allocs = c->buf_misses = 0;
    c->prof_acc.clear();
    return c->prof_text.c_str();
}

int ssn_debug_cycles(ssn_ctx* c, unsigned long long* out64, int reset) {
    return guarded(c, [&] { sync(c); debug_cycles(out64, reset != 0); });
}
int ssn_debug_cycles_persist(ssn_ctx* c, unsigned long long* out256, int reset) {
    return guarded(c, [&] {
        sync(c); debug_cycles_persist(out256, reset != 0);
        unsigned long long z[256]; debug_cycles_dsm(z, reset != 0);       // the cluster kernel of amg_cluster.cu (only one of the two ran)
        for (int i = 0; i < 256; ++i) out256[i] += z[i];
    });
}

int ssn_rng_reset(ssn_ctx* c, uint32_t seed) { return guarded(c, [&] { rng_reset(c, seed); sync(c); }); }
int64_t ssn_rng_drawn(ssn_ctx* c) { return c ? c->rng_drawn : 0; }
int ssn_rand(ssn_ctx* c, int64_t count, double* out) { return guarded(c, [&] { rng_rand(c, count, out); sync(c); }); }

int ssn_malloc(ssn_ctx* c, size_t bytes, void** p) {
    return guarded(c, [&] { SSN_REQUIRE(p, SSN_E_INVALID, "null"); SSN_CUDA(cudaMallocAsync(p, bytes ? bytes : 1, c->stream)); sync(c); });
}
int ssn_free(ssn_ctx* c, void* p) { return guarded(c, [&] { if (p) SSN_CUDA(cudaFreeAsync(p, c->stream)); }); }
int ssn_memcpy_h2d(ssn_ctx* c, void* d, const void* s, size_t b) {
    return guarded(c, [&] { if (b) SSN_CUDA(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, c->stream)); sync(c); });
}
int ssn_memcpy_d2h(ssn_ctx* c, void* d, const void* s, size_t b) {
    return guarded(c, [&] { if (b) SSN_CUDA(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, c->stream)); sync(c); });
}
int ssn_memcpy_d2d(ssn_ctx* c, void* d, const void* s, size_t b) {
    return guarded(c, [&] { if (b) SSN_CUDA(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, c->stream)); sync(c); });
}
int ssn_csr_free(ssn_ctx* c, ssn_csr* A) {
    return guarded(c, [&] {
        if (!A) return;
        if (A->rowptr_dev) SSN_CUDA(cudaFreeAsync(A->rowptr_dev, c->stream));
        if (A->colidx_dev) SSN_CUDA(cudaFreeAsync(A->colidx_dev, c->stream));
        if (A->val_dev) SSN_CUDA(cudaFreeAsync(A->val_dev, c->stream));
        A->rowptr_dev = nullptr; A->colidx_dev = nullptr; A->val_dev = nullptr; A->nnz = 0;
    });
}
int ssn_csr_upload(ssn_ctx* c, int64_t nrows, int64_t ncols, int64_t nnz, const int32_t* rp, const int32_t* ci, const double* v, ssn_csr* out) {
    return guarded(c, [&] {
        SSN_REQUIRE(out && rp && nrows >= 0 && ncols >= 0 && nnz >= 0, SSN_E_INVALID, "csr_upload: bad arguments");
        SSN_REQUIRE(nnz < ((int64_t)1 << 31), SSN_E_TOO_LARGE, "csr_upload: nnz >= 2^31");
        Csr A; A.c = c; A.nrows = nrows; A.ncols = ncols; A.nnz = nnz;
        A.ptr.alloc(c, nrows + 1); A.idx.alloc(c, nnz); A.val.alloc(c, nnz);
        SSN_CUDA(cudaMemcpyAsync(A.ptr.p, rp, sizeof(int) * (nrows + 1), cudaMemcpyHostToDevice, c->stream));
        if (nnz) {
            SSN_CUDA(cudaMemcpyAsync(A.idx.p, ci, sizeof(int) * nnz, cudaMemcpyHostToDevice, c->stream));
            SSN_CUDA(cudaMemcpyAsync(A.val.p, v, sizeof(double) * nnz, cudaMemcpyHostToDevice, c->stream));
        }
        sync(c);
        A.release_to(out);
    });
}
int ssn_csr_download(ssn_ctx* c, const ssn_csr* A, int32_t* rp, int32_t* ci, double* v) {
    return guarded(c, [&] {
        SSN_REQUIRE(A, SSN_E_INVALID, "csr_download: null");
        if (rp) SSN_CUDA(cudaMemcpyAsync(rp, A->rowptr_dev, sizeof(int) * (A->nrows + 1), cudaMemcpyDeviceToHost, c->stream));
        if (ci && A->nnz) SSN_CUDA(cudaMemcpyAsync(ci, A->colidx_dev, sizeof(int) * A->nnz, cudaMemcpyDeviceToHost, c->stream));
        if (v && A->nnz) SSN_CUDA(cudaMemcpyAsync(v, A->val_dev, sizeof(double) * A->nnz, cudaMemcpyDeviceToHost, c->stream));
        sync(c);
    });
}

// ------------------------------------------------------------------ plan operators

int ssn_ax(ssn_ctx* c, const double* x, const double* p, const double* q, int64_t m, int64_t n, double* y) {
    return guarded(c, [&] { plan_ax(c, x, p, q, m, n, y); sync(c); });
}
int ssn_aty(ssn_ctx* c, const double* y, const double* p, const double* q, int64_t m, int64_t n, double* z) {
    return guarded(c, [&] { plan_aty(c, y, p, q, m, n, z); sync(c); });
}

int ssn_ax_host(ssn_ctx* c, const double* x, const double* p, const double* q, int64_t m, int64_t n, double* y) {
    return guarded(c, [&] {
        SSN_REQUIRE(x && p && q && y && m > 0 && n > 0, SSN_E_INVALID, "Ax: bad arguments");
        Buf<double> dx(c, (size_t)m * n), dp(c, m), dq(c, n), dy(c, m + n);
        SSN_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * m * n, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dp.p, p, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        plan_ax(c, dx, dp, dq, m, n, dy);
        SSN_CUDA(cudaMemcpyAsync(y, dy.p, sizeof(double) * (m + n), cudaMemcpyDeviceToHost, c->stream));
        sync(c);
    });
}
int ssn_aty_host(ssn_ctx* c, const double* y, const double* p, const double* q, int64_t m, int64_t n, double* z) {
    return guarded(c, [&] {
        SSN_REQUIRE(z && p && q && y && m > 0 && n > 0, SSN_E_INVALID, "Aty: bad arguments");
        Buf<double> dz(c, (size_t)m * n), dp(c, m), dq(c, n), dy(c, m + n);
        SSN_CUDA(cudaMemcpyAsync(dy.p, y, sizeof(double) * (m + n), cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dp.p, p, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        plan_aty(c, dy, dp, dq, m, n, dz);
        SSN_CUDA(cudaMemcpyAsync(z, dz.p, sizeof(double) * m * n, cudaMemcpyDeviceToHost, c->stream));
        sync(c);
    });
}

int ssn_prox_residual(ssn_ctx* c, const double* w, const double* lam, const double* p, const double* q, int64_t m, int64_t n,
                      double tk, const double* gama, double gama_s, double* axp, double* prox, double* z, uint8_t* s,
                      double* norm2_out, int64_t* count_out) {
    return guarded(c, [&] {
        Buf<double> scal(c, 2);
        plan_prox_residual(c, w, lam, p, q, m, n, tk, gama, gama_s, axp, prox, z, s, scal);
        double h[2]; read_back(c, scal.p, h, 2);
        if (norm2_out) *norm2_out = h[0];
        if (count_out) *count_out = (int64_t)h[1];
    });
}

int ssn_prox_residual_dev(ssn_ctx* c, const double* w, const double* lam, const double* p, const double* q, int64_t m, int64_t n,
                          double tk, const double* gama, double gama_s, double* axp, double* prox, double* z, uint8_t* s, double* scal2_dev) {
    return guarded(c, [&] {
        SSN_REQUIRE(scal2_dev != nullptr, SSN_E_INVALID, "prox_residual_dev: null scalar output");
        plan_prox_residual(c, w, lam, p, q, m, n, tk, gama, gama_s, axp, prox, z, s, scal2_dev);
    });
}

int ssn_prox_residual_pot(ssn_ctx* c, const double* w, const double* lam, const double* p, const double* q, int64_t m, int64_t n,
                          double tk, const double* phi, double* hp, double* prox, uint8_t* s, double* t, double* norm2_out,
                          int64_t* count_out) {
    return guarded(c, [&] {
        Buf<double> scal(c, 3);
        plan_prox_residual_pot(c, w, lam, p, q, m, n, tk, phi, hp, prox, s, t, scal);
        double h[3]; read_back(c, scal.p, h, 3);
        if (norm2_out) *norm2_out = h[0];
        if (count_out) *count_out = (int64_t)h[1];
    });
}

int ssn_prox_trials(ssn_ctx* c, const double* w, const double* lamT, int nt, const double* p, const double* q, int64_t m, int64_t n,
                    double tk, const double* gama, double gama_s, double* n2_out) {
    return guarded(c, [&] { plan_prox_trials(c, w, lamT, nt, p, q, m, n, tk, gama, gama_s, n2_out); sync(c); });
}
int ssn_prox_trials_lin(ssn_ctx* c, const double* w, const double* lam, const double* zeta, const double* p, const double* q,
                        int64_t m, int64_t n, double tk, double delta, int ll0, int nt, double* out) {
    return guarded(c, [&] { plan_prox_trials_lin(c, w, lam, zeta, p, q, m, n, tk, delta, ll0, nt, out, nullptr); sync(c); });
}
int ssn_warmup_class1(ssn_ctx* c, const double* cost, const double* b, const double* p, const double* q, int64_t m, int64_t n,
                      const double* gama, double gama_s, int maxit, double* xk_out, double* lk_out) {
    return guarded(c, [&] { plan_warmup_class1(c, cost, b, p, q, m, n, gama, gama_s, maxit, xk_out, lk_out); sync(c); });
}
int ssn_warm_stage(ssn_ctx* c, int stage, double* xk, double* vk, double* wk, double* pik, double* lk2, double* dd, const double* cost,
                   const double* p, const double* q, const double* b, const double* lk1, const double* axk, const double* y, int64_t m,
                   int64_t n, const double* gama, double gama_s, double ak, double bk, double gk, double* out1, double* out2) {
    return guarded(c, [&] { plan_warm_stage(c, stage, xk, vk, wk, pik, lk2, dd, cost, p, q, b, lk1, axk, y, m, n, gama, gama_s, ak, bk, gk, out1, out2); sync(c); });
}
int ssn_apd_begin(ssn_ctx* c, const double* cost, const double* xk, const double* vk, const double* p, const double* q, int64_t m,
                  int64_t n, double ak, double bk, double* wk_out, double* axk_out) {
    return guarded(c, [&] { plan_apd_begin(c, cost, xk, vk, p, q, m, n, ak, bk, wk_out, axk_out); sync(c); });
}
int ssn_apd_end(ssn_ctx* c, const double* cost, const double* wk, const double* xk, const double* lam, const double* p, const double* q,
                int64_t m, int64_t n, double tk, double ak, const double* gama, double gama_s, double* xk1, double* vk1,
                double* axk1_out, double* cx_out, double* kx2_out) {
    return guarded(c, [&] {
        Buf<double> scal(c, 2);
        plan_apd_end(c, cost, wk, xk, lam, p, q, m, n, tk, ak, gama, gama_s, xk1, vk1, axk1_out, scal);
        double h[2]; read_back(c, scal.p, h, 2);
        if (cx_out) *cx_out = h[0];
        if (kx2_out) *kx2_out = h[1];
    });
}
int ssn_trial_vectors(ssn_ctx* c, const double* lam, const double* zeta, const double* wlk, int64_t N, double delta, int ll0, int nt,
                      double* lamT, double* f0_out) {
    return guarded(c, [&] { plan_trial_vectors(c, lam, zeta, wlk, N, delta, ll0, nt, lamT, f0_out); });
}
int ssn_linesearch(ssn_ctx* c, const double* w, const double* lam_old, const double* zeta, const double* wlk, const double* p,
                   const double* q, int64_t m, int64_t n, double tk, double bk1, const double* gama, double gama_s, double nu,
                   double delta, int ll_max, double cF_old, double ress, int batch, double* lam_new, int* ll_out,
                   double* n2_out, double* cF_out, int* passes_out) {
    return guarded(c, [&] {
        plan_linesearch(c, w, lam_old, zeta, wlk, p, q, m, n, tk, bk1, gama, gama_s, nu, delta, ll_max, cF_old, ress, batch,
                        lam_new, ll_out, n2_out, cF_out, passes_out);
        sync(c);
    });
}

int ssn_apd_ssn_class1(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                       int64_t n, const double* gama, double gama_s, const ssn_apd_options* op, double* xk_out, double* lk_out,
                       ssn_apd_result* res, double* fxk_hist, double* kktx_hist, double* kktl_hist, int32_t* ssn_its_hist,
                       double* steps_host, int64_t steps_cap) {
    return guarded(c, [&] {
        apd_ssn_class1(c, cost, r, l, p, q, m, n, gama, gama_s, op, xk_out, lk_out, res, fxk_hist, kktx_hist, kktl_hist, ssn_its_hist,
                       steps_host, steps_cap);
    });
}
int ssn_apd_ssn_class1_host(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                            int64_t n, const double* gama, double gama_s, const ssn_apd_options* op, double* xk_out, double* lk_out,
                            ssn_apd_result* res, double* fxk_hist, double* kktx_hist, double* kktl_hist, int32_t* ssn_its_hist,
                            double* steps_host, int64_t steps_cap) {
    return guarded(c, [&] {
        SSN_REQUIRE(cost && r && l && p && q && xk_out && lk_out && m > 0 && n > 0, SSN_E_INVALID, "apd_ssn_class1: bad arguments");
        const size_t mn = (size_t)m * n;
        Buf<double> dc(c, mn), dr(c, n), dl(c, m), dp(c, m), dq(c, n), dx(c, mn), dlk(c, m + n), dg;
        SSN_CUDA(cudaMemcpyAsync(dc.p, cost, sizeof(double) * mn, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dr.p, r, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dl.p, l, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dp.p, p, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        if (gama) { dg.alloc(c, mn); SSN_CUDA(cudaMemcpyAsync(dg.p, gama, sizeof(double) * mn, cudaMemcpyHostToDevice, c->stream)); }
        apd_ssn_class1(c, dc, dr, dl, dp, dq, m, n, gama ? dg.p : nullptr, gama_s, op, dx, dlk, res, fxk_hist, kktx_hist, kktl_hist,
                       ssn_its_hist, steps_host, steps_cap);
        SSN_CUDA(cudaMemcpyAsync(xk_out, dx.p, sizeof(double) * mn, cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaMemcpyAsync(lk_out, dlk.p, sizeof(double) * (m + n), cudaMemcpyDeviceToHost, c->stream));
        sync(c);
    });
}

int ssn_ssn_step_class1(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                        int64_t n, double bk1, double tk, const double* gama, double gama_s, int inner_solver, const ssn_amg_options* amg,
                        double* lk_new, double* Fk_new, double* info12) {
    return guarded(c, [&] { ssn_step_class1(c, wk, lk, wlk, p, q, m, n, bk1, tk, gama, gama_s, inner_solver, amg, lk_new, Fk_new, info12); sync(c); });
}
int ssn_ssn_step_class1_host(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                             int64_t n, double bk1, double tk, const double* gama, double gama_s, int inner_solver,
                             const ssn_amg_options* amg, double* lk_new, double* Fk_new, double* info12) {
    return guarded(c, [&] {
        SSN_REQUIRE(wk && lk && wlk && p && q && lk_new && Fk_new && m > 0 && n > 0, SSN_E_INVALID, "ssn_step_class1: bad arguments");
        const size_t mn = (size_t)m * n, N = (size_t)(m + n);
        Buf<double> dw(c, mn), dlk(c, N), dwl(c, N), dp(c, m), dq(c, n), dln(c, N), dF(c, N), dg;
        SSN_CUDA(cudaMemcpyAsync(dw.p, wk, sizeof(double) * mn, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dlk.p, lk, sizeof(double) * N, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dwl.p, wlk, sizeof(double) * N, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dp.p, p, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        if (gama) { dg.alloc(c, mn); SSN_CUDA(cudaMemcpyAsync(dg.p, gama, sizeof(double) * mn, cudaMemcpyHostToDevice, c->stream)); }
        ssn_step_class1(c, dw, dlk, dwl, dp, dq, m, n, bk1, tk, gama ? dg.p : nullptr, gama_s, inner_solver, amg, dln, dF, info12);
        SSN_CUDA(cudaMemcpyAsync(lk_new, dln.p, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaMemcpyAsync(Fk_new, dF.p, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        sync(c);
    });
}

// ---- Class 2 (partial OT): the script, its warm start and one SsN step as single calls
int ssn_warmup_class2(ssn_ctx* c, const double* cost, const double* b, const double* p, const double* q, int64_t m, int64_t n,
                      const double* phi, int maxit, double* uk_out, double* lk_out) {
    return guarded(c, [&] { plan_warmup_class2(c, cost, b, p, q, m, n, phi, maxit, uk_out, lk_out); sync(c); });
}
int ssn_apd_begin_pot(ssn_ctx* c, const double* cost, const double* uk, const double* vk, const double* p, const double* q, int64_t m,
                      int64_t n, const double* phi, const double* b, const double* lk, double ak, double bk, double bk1, double* wk_out,
                      double* huk_out, double* wlk_out) {
    return guarded(c, [&] { plan_apd_begin_pot(c, cost, uk, vk, p, q, m, n, phi, b, lk, ak, bk, bk1, wk_out, huk_out, wlk_out); sync(c); });
}
int ssn_apd_end_pot(ssn_ctx* c, const double* cost, const double* wk, const double* uk, const double* lk, const double* p, const double* q,
                    int64_t m, int64_t n, const double* phi, const double* b, double tk, double ak, double* uk1, double* vk1,
                    double* huk1_out, double* scal5_host) {
    return guarded(c, [&] {
        Buf<double> scal(c, 5);
        plan_apd_end_pot(c, cost, wk, uk, lk, p, q, m, n, phi, b, tk, ak, uk1, vk1, huk1_out, scal);
        double h[5]; read_back(c, scal.p, h, 5);
        if (scal5_host) for (int i = 0; i < 5; ++i) scal5_host[i] = h[i];
    });
}
int ssn_apd_ssn_class2(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                       int64_t n, double mu, const double* phi, const ssn_apd_options* op, double* uk_out, double* lk_out,
                       ssn_apd_result* res, double* fxk_hist, double* kkt4_hist, int32_t* ssn_its_hist, double* steps_host,
                       int64_t steps_cap) {
    return guarded(c, [&] {
        apd_ssn_class2(c, cost, r, l, p, q, m, n, mu, phi, op, uk_out, lk_out, res, fxk_hist, kkt4_hist, ssn_its_hist, steps_host, steps_cap);
    });
}
int ssn_apd_ssn_class2_host(ssn_ctx* c, const double* cost, const double* r, const double* l, const double* p, const double* q, int64_t m,
                            int64_t n, double mu, const double* phi, const ssn_apd_options* op, double* uk_out, double* lk_out,
                            ssn_apd_result* res, double* fxk_hist, double* kkt4_hist, int32_t* ssn_its_hist, double* steps_host,
                            int64_t steps_cap) {
    return guarded(c, [&] {
        SSN_REQUIRE(cost && r && l && p && q && phi && uk_out && lk_out && m > 0 && n > 0, SSN_E_INVALID, "apd_ssn_class2: bad arguments");
        const size_t mn = (size_t)m * n, N = (size_t)(m + n);
        Buf<double> dc(c, mn), dphi(c, mn), dr(c, n), dl(c, m), dp(c, m), dq(c, n), du(c, mn + N), dlk(c, N + 1);
        SSN_CUDA(cudaMemcpyAsync(dc.p, cost, sizeof(double) * mn, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dphi.p, phi, sizeof(double) * mn, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dr.p, r, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dl.p, l, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dp.p, p, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        apd_ssn_class2(c, dc, dr, dl, dp, dq, m, n, mu, dphi, op, du, dlk, res, fxk_hist, kkt4_hist, ssn_its_hist, steps_host, steps_cap);
        SSN_CUDA(cudaMemcpyAsync(uk_out, du.p, sizeof(double) * (mn + N), cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaMemcpyAsync(lk_out, dlk.p, sizeof(double) * (N + 1), cudaMemcpyDeviceToHost, c->stream));
        sync(c);
    });
}
int ssn_ssn_step_class2(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                        int64_t n, double bk1, double tk, const double* phi, int inner_solver, const ssn_amg_options* amg,
                        const ssn_pcg_options* pcg, double* lk_new, double* Fk_new, double* info12) {
    return guarded(c, [&] { ssn_step_class2(c, wk, lk, wlk, p, q, m, n, bk1, tk, phi, inner_solver, amg, pcg, lk_new, Fk_new, info12); sync(c); });
}
int ssn_ssn_step_class2_host(ssn_ctx* c, const double* wk, const double* lk, const double* wlk, const double* p, const double* q, int64_t m,
                             int64_t n, double bk1, double tk, const double* phi, int inner_solver, const ssn_amg_options* amg,
                             const ssn_pcg_options* pcg, double* lk_new, double* Fk_new, double* info12) {
    return guarded(c, [&] {
        SSN_REQUIRE(wk && lk && wlk && p && q && phi && lk_new && Fk_new && m > 0 && n > 0, SSN_E_INVALID, "ssn_step_class2: bad arguments");
        const size_t mn = (size_t)m * n, N = (size_t)(m + n);
        Buf<double> dw(c, mn + N), dphi(c, mn), dlk(c, N + 1), dwl(c, N + 1), dp(c, m), dq(c, n), dln(c, N + 1), dF(c, N + 1);
        SSN_CUDA(cudaMemcpyAsync(dw.p, wk, sizeof(double) * (mn + N), cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dphi.p, phi, sizeof(double) * mn, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dlk.p, lk, sizeof(double) * (N + 1), cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dwl.p, wlk, sizeof(double) * (N + 1), cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dp.p, p, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        ssn_step_class2(c, dw, dlk, dwl, dp, dq, m, n, bk1, tk, dphi, inner_solver, amg, pcg, dln, dF, info12);
        SSN_CUDA(cudaMemcpyAsync(lk_new, dln.p, sizeof(double) * (N + 1), cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaMemcpyAsync(Fk_new, dF.p, sizeof(double) * (N + 1), cudaMemcpyDeviceToHost, c->stream));
        sync(c);
    });
}
int ssn_amg4pot_str(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, int twogrid, double* zeta, int* it, double* res, int* info) {
    return guarded(c, [&] { amg4pot(c, pd, opts, zeta, it, res, info, twogrid != 0); });
}

int ssn_asat(ssn_ctx* c, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n, ssn_csr* H) {
    return guarded(c, [&] { SSN_REQUIRE(H, SSN_E_INVALID, "ASAt: null output"); Csr h = asat(c, s, p, q, m, n); sync(c); h.release_to(H); });
}
int ssn_asat_host(ssn_ctx* c, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n, ssn_csr* H) {
    return guarded(c, [&] {
        SSN_REQUIRE(H && s && p && q && m > 0 && n > 0, SSN_E_INVALID, "ASAt: bad arguments");
        Buf<uint8_t> ds(c, (size_t)m * n); Buf<double> dp(c, m), dq(c, n);
        SSN_CUDA(cudaMemcpyAsync(ds.p, s, (size_t)m * n, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dp.p, p, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        Csr h = asat(c, ds, dp, dq, m, n); sync(c); h.release_to(H);
    });
}
int ssn_asat_coo(ssn_ctx* c, const int64_t* lin, int64_t E, const double* p, const double* q, int64_t m, int64_t n, ssn_csr* H) {
    return guarded(c, [&] { SSN_REQUIRE(H, SSN_E_INVALID, "ASAt: null output"); Csr h = asat_coo(c, (const long long*)lin, E, p, q, m, n); sync(c); h.release_to(H); });
}
int ssn_active_coo(ssn_ctx* c, const uint8_t* s, int64_t m_loc, int64_t n, int64_t row_offset, int64_t m_global, int64_t** lin_out, int64_t* E_out) {
    return guarded(c, [&] {
        SSN_REQUIRE(s && lin_out && E_out, SSN_E_INVALID, "active_coo: null");
        long long* l = nullptr; int64_t E = 0;
        active_coo(c, s, m_loc, n, row_offset, m_global, &l, &E); sync(c);
        *lin_out = (int64_t*)l; *E_out = E;
    });
}
int ssn_asatz(ssn_ctx* c, const double* z, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n, double* y) {
    return guarded(c, [&] { asatz(c, z, s, p, q, m, n, y); sync(c); });
}
int ssn_invaat(ssn_ctx* c, const double* x, const double* p, const double* q, int64_t m, int64_t n, double sg1, double sg2, double* y) {
    return guarded(c, [&] { invaat(c, x, p, q, m, n, sg1, sg2, y); sync(c); });
}
int ssn_invhht(ssn_ctx* c, const double* v, const double* p, const double* q, int64_t m, int64_t n, double sg, const double* phi, double* y) {
    return guarded(c, [&] { invhht(c, v, p, q, m, n, sg, phi, y); sync(c); });
}

// ------------------------------------------------------------------ AMG setup

int ssn_strength(ssn_ctx* c, const ssn_csr* A, int which, ssn_csr* S) {
    return guarded(c, [&] { SSN_REQUIRE(A && S, SSN_E_INVALID, "strength: null"); Csr s = strength_matrix(c, *A, which == 1 ? 1 : 2); sync(c); s.release_to(S); });
}
int ssn_mis_set(ssn_ctx* c, const ssn_csr* A, double theta, uint8_t* isC, uint8_t* isF, ssn_csr* As) {
    return guarded(c, [&] {
        SSN_REQUIRE(A && isC && isF, SSN_E_INVALID, "mis_set: null");
        Buf<uint8_t> flags;
        mis_set(c, *A, theta, isC, isF, flags);
        if (As) { Csr s = flags_to_csr(c, *A, flags); sync(c); s.release_to(As); }
        sync(c);
    });
}
int ssn_cf_split(ssn_ctx* c, const ssn_csr* S, uint8_t* indC, uint8_t* indF) {
    return guarded(c, [&] { SSN_REQUIRE(S && indC && indF, SSN_E_INVALID, "cf_split: null"); cf_split(c, *S, indC, indF); sync(c); });
}
int ssn_transfer(ssn_ctx* c, const ssn_csr* A, const ssn_amg_options* opts, int level_J, ssn_csr* Ac, ssn_csr* Pro, ssn_csr* As, uint8_t* indC) {
    return guarded(c, [&] {
        SSN_REQUIRE(A && Ac && Pro, SSN_E_INVALID, "transfer: null");
        AmgOptions o;
        if (opts) o = resolve_options(opts);
        else { o = resolve_options(nullptr); o.theta = 1.0 / 40; o.bigph = 0; o.inter = 1; o.isnsp = 0; }   // transfer.m:8-15
        Csr ac, pro; Buf<uint8_t> isC, flags;
        transfer(c, *A, o, level_J, ac, pro, &isC, As ? &flags : nullptr);
        if (indC) SSN_CUDA(cudaMemcpyAsync(indC, isC.p, A->nrows, cudaMemcpyDeviceToDevice, c->stream));
        if (As) { Csr s = flags_to_csr(c, *A, flags); sync(c); s.release_to(As); }
        sync(c);
        ac.release_to(Ac); pro.release_to(Pro);
    });
}
int ssn_amg_setup(ssn_ctx* c, const ssn_csr* A, const ssn_amg_options* opts, int* levels) {
    return guarded(c, [&] {
        SSN_REQUIRE(A, SSN_E_INVALID, "amg_setup: null");
        amg_setup(c, *A, resolve_options(opts)); sync(c);
        if (levels) *levels = c->hier->J;
    });
}
int ssn_amg_level(ssn_ctx* c, int k, ssn_csr* A, ssn_csr* Pro) {
    return guarded(c, [&] {
        SSN_REQUIRE(c->hier, SSN_E_NO_HIERARCHY, "no live hierarchy");
        SSN_REQUIRE(k >= 1 && k <= c->hier->J, SSN_E_INVALID, "level out of range");
        Level& L = c->hier->lv[k - 1];
        if (A) *A = L.A.view();
        if (Pro) { if (k > 1) *Pro = L.P.view(); else std::memset(Pro, 0, sizeof(*Pro)); }
    });
}
int ssn_amg_clear(ssn_ctx* c) { return guarded(c, [&] { amg_clear(c); }); }

int ssn_mg_vcycle(ssn_ctx* c, const double* r, int isnsp, int k, double* e) {
    return guarded(c, [&] { SSN_REQUIRE(r && e, SSN_E_INVALID, "null"); mg_cycle(c, r, isnsp, k, e, false, true); sync(c); });
}
int ssn_mg_wcycle(ssn_ctx* c, const double* r, int isnsp, int k, double* e) {
    return guarded(c, [&] { SSN_REQUIRE(r && e, SSN_E_INVALID, "null"); mg_cycle(c, r, isnsp, k, e, true, false); sync(c); });
}
int ssn_class_amg(ssn_ctx* c, const ssn_csr* A, const double* b, const ssn_amg_options* opts, int keep, double* x, int* it,
                  double* rel_res, double* rel_resk, double* rhok, int* hist_len) {
    return guarded(c, [&] {
        SSN_REQUIRE(A && b && x, SSN_E_INVALID, "Class_AMG: null");
        class_amg(c, *A, b, resolve_options(opts), keep != 0, x, it, rel_res, rel_resk, rhok, hist_len); sync(c);
    });
}
int ssn_pcg(ssn_ctx* c, const ssn_csr* H, const double* e, const ssn_pcg_options* opts, double* d, int* it, double* res, double* resk) {
    return guarded(c, [&] { SSN_REQUIRE(H && e && d, SSN_E_INVALID, "PCG: null"); pcg_solve(c, *H, e, opts, d, it, res, resk); sync(c); });
}

// ------------------------------------------------------------------ dispatch

int ssn_components(ssn_ctx* c, const ssn_csr* A, int32_t* blocks, int32_t* sizes, int32_t* p, int32_t* r, int* ncomp) {
    return guarded(c, [&] {
        SSN_REQUIRE(A && blocks && sizes && p && r, SSN_E_INVALID, "components: null");
        components(c, *A, blocks, sizes, p, r, ncomp); sync(c);
    });
}
int ssn_hybrid_amg(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, double* zeta, int* it, double* res, int* info) {
    return guarded(c, [&] { hybrid_amg(c, pd, opts, zeta, it, res, info); });
}
int ssn_hybrid_twogrid(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, double* zeta, int* it, double* res, int* info) {
    return guarded(c, [&] { hybrid_amg(c, pd, opts, zeta, it, res, info, true); });
}
int ssn_twogrid_bigph(ssn_ctx* c, const ssn_csr* A, const double* b, const ssn_amg_options* opts, double* x, int* it,
                      double* rel_res, double* rel_resk, double* rhok, int* hist_len) {
    return guarded(c, [&] {
        SSN_REQUIRE(A && b && x, SSN_E_INVALID, "twogrid_bigph: null");
        twogrid_bigph(c, *A, b, resolve_options(opts), x, it, rel_res, rel_resk, rhok, hist_len); sync(c);
    });
}
int ssn_twogrid(ssn_ctx* c, const ssn_csr* A, const double* b, const ssn_amg_options* opts, double* x, int* it,
                double* rel_res, double* rel_resk, double* rhok, int* hist_len) {
    return guarded(c, [&] {
        SSN_REQUIRE(A && b && x, SSN_E_INVALID, "twogrid: null");
        twogrid_bigph(c, *A, b, resolve_options(opts), x, it, rel_res, rel_resk, rhok, hist_len, true); sync(c);
    });
}
int ssn_aug_pcg(ssn_ctx* c, const ssn_prob_data* pd, const ssn_pcg_options* opts, double* zeta, int* it, double* res, int* info) {
    return guarded(c, [&] { aug_pcg(c, pd, opts, zeta, it, res, info); });
}
int ssn_amg4pot(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, double* zeta, int* it, double* res, int* info) {
    return guarded(c, [&] { amg4pot(c, pd, opts, zeta, it, res, info); });
}
int ssn_pcg4pot(ssn_ctx* c, const ssn_prob_data* pd, const ssn_pcg_options* opts, double* zeta, int* it, double* res, int* info) {
    return guarded(c, [&] { pcg4pot(c, pd, opts, zeta, it, res, info); });
}
int ssn_rescaled_system(ssn_ctx* c, const ssn_prob_data* pd, ssn_csr* Ae, double* f) {
    return guarded(c, [&] {
        SSN_REQUIRE(Ae, SSN_E_INVALID, "rescaled_system: null");
        Csr ae; Buf<double> qp, Kd;
        rescaled_system(c, pd, ae, f, qp, Kd); sync(c); ae.release_to(Ae);
    });
}

int ssn_jk_system(ssn_ctx* c, const ssn_prob_data* pd, ssn_csr* Jk) {
    return guarded(c, [&] {
        SSN_REQUIRE(Jk, SSN_E_INVALID, "jk_system: null");
        Csr jk; jk_system(c, pd, jk); sync(c); jk.release_to(Jk);
    });
}

// ------------------------------------------------------------------ sparse utilities

int ssn_spmv(ssn_ctx* c, const ssn_csr* A, const double* x, double* y) {
    return guarded(c, [&] { SSN_REQUIRE(A && x && y, SSN_E_INVALID, "spmv: null"); spmv(c, *A, x, y); sync(c); });
}
int ssn_spgemm(ssn_ctx* c, const ssn_csr* A, const ssn_csr* B, ssn_csr* C) {
    return guarded(c, [&] { SSN_REQUIRE(A && B && C, SSN_E_INVALID, "spgemm: null"); Csr r = spgemm(c, *A, *B); sync(c); r.release_to(C); });
}
int ssn_transpose(ssn_ctx* c, const ssn_csr* A, ssn_csr* At) {
    return guarded(c, [&] { SSN_REQUIRE(A && At, SSN_E_INVALID, "transpose: null"); Csr r = transpose(c, *A); sync(c); r.release_to(At); });
}

}  // extern "C"
