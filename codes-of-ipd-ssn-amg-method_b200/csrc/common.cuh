// common.cuh -- context, error handling, stream-ordered device buffers and small device
// helpers shared by every translation unit of libssnamg.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <initializer_list>
#include <string>
#include <vector>
#include <memory>
#include <stdexcept>
#include <map>
#include <unordered_map>
#include <chrono>

#include "../../include/ssnamg.h"

namespace ssn {

struct Error : std::exception {
    int code; std::string msg;
    Error(int c, std::string m) : code(c), msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

#define SSN_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            throw ::ssn::Error(SSN_E_CUDA, std::string(#call) + ": " +                     \
                               cudaGetErrorString(e__) + " (" + __FILE__ + ":" +           \
                               std::to_string(__LINE__) + ")");                            \
    } while (0)

#define SSN_REQUIRE(cond, code, text)                                                      \
    do { if (!(cond)) throw ::ssn::Error((code), (text)); } while (0)

struct Hierarchy;   // amg.cuh

}  // namespace ssn

// The opaque context of the C ABI.
struct ssn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int num_sms = 148;
    size_t smem_optin = 0;
    // pinned host scratch for scalar read-backs
    double* h_pin = nullptr;              // 4096 doubles
    static constexpr int kPinDoubles = 4096;
    // small device->host reads without a copy-engine transfer and a stream synchronise: a one-block kernel stores the values
    // into MAPPED pinned memory, then a sequence number; the host spins on the number (SSN_POLL_READS=0: memcpy + synchronise)
    bool poll_reads = true;
    double* h_poll = nullptr;             // kPinDoubles doubles + 1 flag word, cudaHostAlloc(mapped)
    double* d_poll = nullptr;             // the same memory as the device sees it
    unsigned long long poll_seq = 0;
    // small host->device uploads (level tables, lists) go through a ring of PINNED slots: cudaMemcpyAsync from pageable
    // memory synchronises the stream before it copies, i.e. every such upload was a host wait for all queued kernels
    static constexpr int kUpSlots = 32, kUpBytes = 4096;
    unsigned char* h_up = nullptr;        // kUpSlots * kUpBytes bytes, cudaMallocHost
    int up_next = 0;
    // MATLAB random stream (device-resident MT19937 state)
    uint32_t* mt_state = nullptr;         // 624 words + 1 index word
    int64_t rng_drawn = 0;
    // Class_AMG hierarchy (the reference's globals Ack/Prok/Rk/J/smoth_it)
    ssn::Hierarchy* hier = nullptr;
    // optional phase profiler (ssn_profile_enable): wall time per named phase, stream-synchronised
    bool no_cluster = true;               // SSN_CLUSTER=1 enables the 8-CTA cluster cycle kernel
    int64_t persist_max_nnz = (int64_t)1 << 40;   // SSN_PERSIST_MAXNNZ: above this the cycle is launched kernel by kernel
    bool persist = true;                  // SSN_PERSIST=0: launch the large-level cycle kernel by kernel
    bool fused_setup = false;             // SSN_FUSED_SETUP=1: the small levels of the hierarchy are coarsened by ONE kernel (one CTA; same
                                          // hierarchy bit for bit, but slower than kernel by kernel on a B200: opt-in)
    bool cluster_solve = true;            // SSN_CLUSTER_SOLVE=0: the persistent solve always runs grid-wide (cooperative launch)
    static constexpr int kSpgemmSites = 64;
    int last_dsm_halo = -1;               // halo entries of CTA 0 in the last dsm_solve_kernel launch (diagnostic)
    int64_t spgemm_slab_limit = (int64_t)1 << 30;   // sparse products with more intermediate entries are formed in slabs of rows
    unsigned spgemm_epoch = 0;
    int spgemm_site = -1;                 // >= 0 inside amg_setup: index of the next sparse product of this hierarchy
    unsigned char spgemm_big[kSpgemmSites] = {0};   // 1: the optimistic warp-path attempt of that product failed last time
    bool stage_dense = true;              // SSN_STAGE_DENSE=0: the grid-wide solve kernel gathers from L2 on dense levels too
    bool mis_cluster = true;              // SSN_MIS_CLUSTER=0: the MIS rounds of mis_set.m launch by launch (one host read per round)
    bool dsm_solve = true;                // SSN_DSM_SOLVE=0: the cluster solve keeps its vectors in global memory (first cluster kernel)
    int64_t cluster_max_nnz = (int64_t)1 << 20;   // SSN_CLUSTER_MAXNNZ: larger hierarchies (explicit levels) use the grid-wide kernel
    bool dense_tail = true;               // SSN_DENSE_TAIL=0 falls back to the step-by-step tail kernel
    int ls_max_nt = 128;                  // SSN_LS_MAXNT: largest batch of the screened line search (8..128 steps per read of w)
    int plan_waves = 0;                   // SSN_PLAN_WAVES: waves of blocks of the plan-wide reduction kernels (0: two)
    bool tg_cluster = true;               // SSN_TG_CLUSTER=0: twogrid_bigph's iteration loop kernel by kernel (coarse PCG: the grid-wide pcg_kernel)
    bool plan_stage = true;               // SSN_PLAN_STAGE=0: the plan-wide reduction kernels load straight into registers (no cp.async staging)
    bool ls_screen = true;                // SSN_LS_SCREEN=0: the adaptive line search uses the dense 8-trial kernel only
    double ls_last_density = -1.0;        // share of the plan's entries that survived the screen in the last screened batch (< 0: none yet)
    int ls_last_ll = -1;                  // backtracking steps the last line search of this context accepted at (< 0: none yet)
    bool device_setup = true;             // SSN_DEVICE_SETUP=0: SSOR / IC(0) factors and their levels built on the host instead of the device (trifactor.cu)
    int small_scan_max = 1 << 14;         // SSN_SMALL_SCAN_MAX: largest array scanned by the one-block kernel (cub::DeviceScan above)
    int dense_max_n = 2048;               // SSN_DENSE_MAXN: largest level collapsed into a dense operator
    // CUDA-event timer around the launches of the plan-wide kernels (ssn_kernel_timer): bench.py's roofline
    bool ktimer = false; cudaEvent_t kt0 = nullptr, kt1 = nullptr; double kt_ms = 0.0; int64_t kt_n = 0;
    bool prof = false;
    // Recycled device buffers (SSN_BUF_CACHE=0: every Buf is a cudaMallocAsync / cudaFreeAsync pair).  A hierarchy setup makes
    // several hundred short-lived buffers of a few KB; all work of a context is ordered on ONE stream, so a freed block can be
    // handed to the next request of its size class without a runtime call.  Blocks of up to 8 MB, power-of-two classes.
    static constexpr int kCacheMinShift = 9, kCacheMaxShift = 23;
    bool buf_cache = true;
    std::vector<void*> cache_free[kCacheMaxShift + 1];
    std::unordered_map<void*, unsigned char> cache_live;      // blocks handed out by the cache -> size class
    size_t cache_bytes = 0, cache_cap = (size_t)512 << 20;    // bytes parked in the free lists, and their bound
    int64_t buf_allocs = 0, buf_misses = 0;                   // requests / requests that went to cudaMallocAsync
    std::map<std::string, std::pair<double, long>> prof_acc;
    std::string prof_text;
};

namespace ssn {

// ---- the context's block cache (see ssn_ctx::buf_cache)
inline void* ctx_alloc(ssn_ctx* c, size_t bytes) {
    ++c->buf_allocs;
    void* p = nullptr;
    if (c->buf_cache && bytes <= ((size_t)1 << ssn_ctx::kCacheMaxShift)) {
        int cls = ssn_ctx::kCacheMinShift;
        while (((size_t)1 << cls) < bytes) ++cls;
        auto& fl = c->cache_free[cls];
        if (!fl.empty()) { p = fl.back(); fl.pop_back(); c->cache_bytes -= (size_t)1 << cls; }
        else {
            ++c->buf_misses;
            cudaError_t e = cudaMallocAsync(&p, (size_t)1 << cls, c->stream);
            if (e != cudaSuccess) throw Error(SSN_E_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        }
        c->cache_live[p] = (unsigned char)cls;
        return p;
    }
    ++c->buf_misses;
    cudaError_t e = cudaMallocAsync(&p, bytes, c->stream);
    if (e != cudaSuccess) throw Error(SSN_E_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    return p;
}
inline void ctx_free(ssn_ctx* c, void* p) {
    if (!p) return;
    auto it = c->cache_live.find(p);
    if (it == c->cache_live.end()) { cudaFreeAsync(p, c->stream); return; }
    const int cls = it->second;
    c->cache_live.erase(it);
    if (c->cache_bytes + ((size_t)1 << cls) <= c->cache_cap) { c->cache_free[cls].push_back(p); c->cache_bytes += (size_t)1 << cls; }
    else cudaFreeAsync(p, c->stream);
}
// the block leaves the cache's books: its new owner frees it with cudaFreeAsync (ssn_free, ssn_csr_free)
inline void ctx_untrack(ssn_ctx* c, void* p) { if (p) c->cache_live.erase(p); }
inline void ctx_cache_purge(ssn_ctx* c) {
    for (auto& fl : c->cache_free) { for (void* p : fl) cudaFreeAsync(p, c->stream); fl.clear(); }
    c->cache_bytes = 0;
}

// Stream-ordered device buffer (the context's block cache over the cudaMallocAsync pool; frees are stream-ordered too).
template <class T>
struct Buf {
    ssn_ctx* c = nullptr; T* p = nullptr; size_t n = 0;
    bool owned = true;                    // false: a view into memory somebody else owns (an arena of the fused setup kernel)
    Buf() = default;
    Buf(ssn_ctx* ctx, size_t count) { alloc(ctx, count); }
    Buf(const Buf&) = delete; Buf& operator=(const Buf&) = delete;
    Buf(Buf&& o) noexcept : c(o.c), p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; }
    Buf& operator=(Buf&& o) noexcept {
        if (this != &o) { reset(); c = o.c; p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~Buf() { reset(); }
    static Buf view(ssn_ctx* ctx, T* ptr, size_t count) { Buf b; b.c = ctx; b.p = ptr; b.n = count; b.owned = false; return b; }
    void alloc(ssn_ctx* ctx, size_t count) {
        reset(); c = ctx; n = count; owned = true;
        size_t bytes = (count ? count : 1) * sizeof(T);
        p = static_cast<T*>(ctx_alloc(ctx, bytes));
    }
    void reset() {
        if (p && owned) ctx_free(c, p);
        p = nullptr; n = 0; owned = true;
    }
    // hands the pointer to a caller that will cudaFreeAsync it: a view is copied into an owned allocation first
    T* release() {
        if (p && !owned) {
            T* q = nullptr;
            SSN_CUDA(cudaMallocAsync((void**)&q, (n ? n : 1) * sizeof(T), c->stream));
            if (n) SSN_CUDA(cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
            p = nullptr; n = 0; owned = true;
            return q;
        }
        T* r = p; p = nullptr; n = 0; ctx_untrack(c, r); return r;
    }
    void zero() { SSN_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), c->stream)); }
    operator T*() const { return p; }
    T* get() const { return p; }
};

// Owning device CSR (internal); converts to / from the ABI's ssn_csr.
struct Csr {
    ssn_ctx* c = nullptr;
    int64_t nrows = 0, ncols = 0, nnz = 0;
    Buf<int> ptr; Buf<int> idx; Buf<double> val;
    Csr() = default;
    Csr(Csr&&) = default; Csr& operator=(Csr&&) = default;
    ssn_csr view() const {
        ssn_csr v; v.nrows = nrows; v.ncols = ncols; v.nnz = nnz;
        v.rowptr_dev = ptr.p; v.colidx_dev = idx.p; v.val_dev = val.p; return v;
    }
    // hand the arrays over to the ABI struct (caller frees with ssn_csr_free)
    void release_to(ssn_csr* out) {
        out->nrows = nrows; out->ncols = ncols; out->nnz = nnz;
        out->rowptr_dev = ptr.release(); out->colidx_dev = idx.release(); out->val_dev = val.release();
    }
};

// Non-owning view used by kernels and internal functions.
struct CsrView {
    int nrows = 0, ncols = 0; int64_t nnz = 0;
    const int* ptr = nullptr; const int* idx = nullptr; const double* val = nullptr;
    CsrView() = default;
    CsrView(const ssn_csr& a) : nrows((int)a.nrows), ncols((int)a.ncols), nnz(a.nnz),
        ptr(a.rowptr_dev), idx(a.colidx_dev), val(a.val_dev) {}
    CsrView(const Csr& a) : nrows((int)a.nrows), ncols((int)a.ncols), nnz(a.nnz),
        ptr(a.ptr.p), idx(a.idx.p), val(a.val.p) {}
};

inline void check_launch(ssn_ctx* c, const char* what) {
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw Error(SSN_E_CUDA, std::string("launch ") + what + ": " + cudaGetErrorString(e));
}

#define SSN_LAUNCH(ctx, kernel, grid, block, smem, ...)                                    \
    do {                                                                                   \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                   \
        ::ssn::check_launch((ctx), #kernel);                                               \
    } while (0)

// Times ONE kernel launch with CUDA events on the launching stream when ssn_kernel_timer is on.
struct KernelTimer {
    ssn_ctx* c;
    bool armed = false;
    explicit KernelTimer(ssn_ctx* ctx) : c(ctx) {
        if (c->ktimer) {
            if (!c->kt0) { if (cudaEventCreate(&c->kt0) != cudaSuccess || cudaEventCreate(&c->kt1) != cudaSuccess) { c->kt0 = c->kt1 = nullptr; note("cudaEventCreate"); return; } }
            armed = cudaEventRecord(c->kt0, c->stream) == cudaSuccess;
            if (!armed) note("cudaEventRecord(start)");
        }
    }
    ~KernelTimer() {
        if (c->ktimer && armed) {
            float ms = 0.f;
            if (cudaEventRecord(c->kt1, c->stream) != cudaSuccess) { note("cudaEventRecord(stop)"); return; }
            if (cudaEventSynchronize(c->kt1) != cudaSuccess) { note("cudaEventSynchronize"); return; }
            if (cudaEventElapsedTime(&ms, c->kt0, c->kt1) != cudaSuccess) { note("cudaEventElapsedTime"); return; }
            c->kt_ms += ms; c->kt_n += 1;
        }
    }
    void note(const char* what) { const cudaError_t e = cudaGetLastError(); c->err = std::string("kernel timer: ") + what + ": " + cudaGetErrorString(e); fprintf(stderr, "%s\n", c->err.c_str()); }
};

struct Phase {
    ssn_ctx* c; const char* name; std::chrono::steady_clock::time_point t0; long l0;
    Phase(ssn_ctx* ctx, const char* n) : c(ctx), name(n) {
        if (c->prof) { cudaStreamSynchronize(c->stream); t0 = std::chrono::steady_clock::now(); l0 = c->launches; }
    }
    ~Phase() {
        if (c->prof) {
            cudaStreamSynchronize(c->stream);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            auto& a = c->prof_acc[name]; a.first += ms; a.second += 1;
            auto& b = c->prof_acc[std::string(name) + " #launches"]; b.first += (double)(c->launches - l0); b.second += 1;
        }
    }
};

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// values of up to 4 device ints / one device array -> the context's mapped pinned buffer, visible to the host when the
// sequence word changes (sparse.cu); returns after the values have arrived
void poll_read(ssn_ctx* c, const void* dev, size_t bytes);
void poll_read_ints(ssn_ctx* c, const int* const* src, int k);

// asynchronous upload of a few KB from any host memory (staged through the pinned ring; larger blocks: plain copy)
inline void upload_small(ssn_ctx* c, void* dst_dev, const void* src_host, size_t bytes) {
    if (bytes == 0) return;
    if (c->h_up && bytes <= (size_t)ssn_ctx::kUpBytes) {
        unsigned char* slot = c->h_up + (size_t)c->up_next * ssn_ctx::kUpBytes;
        c->up_next = (c->up_next + 1) % ssn_ctx::kUpSlots;
        std::memcpy(slot, src_host, bytes);
        SSN_CUDA(cudaMemcpyAsync(dst_dev, slot, bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        SSN_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, c->stream));
    }
}

// read `count` elements (<= pinned scratch) back to the host, synchronously
template <class T>
inline void read_back(ssn_ctx* c, const T* dev, T* host, size_t count) {
    size_t bytes = count * sizeof(T);
    if (c->poll_reads && c->h_poll && bytes > 0 && bytes <= sizeof(double) * ssn_ctx::kPinDoubles && bytes % 4 == 0) {
        poll_read(c, dev, bytes);
        std::memcpy(host, c->h_poll, bytes);
        return;
    }
    if (bytes <= sizeof(double) * ssn_ctx::kPinDoubles) {
        SSN_CUDA(cudaMemcpyAsync(c->h_pin, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaStreamSynchronize(c->stream));
        std::memcpy(host, c->h_pin, bytes);
    } else {
        SSN_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        SSN_CUDA(cudaStreamSynchronize(c->stream));
    }
}
template <class T>
inline T read_scalar(ssn_ctx* c, const T* dev) { T v; read_back(c, dev, &v, 1); return v; }
// several device ints with ONE stream synchronisation (the copies queue up behind the kernels that produce them)
inline void read_ints(ssn_ctx* c, std::initializer_list<const int*> src, int* out) {
    if (c->poll_reads && c->h_poll && src.size() <= 4) {
        const int* ptrs[4] = {nullptr, nullptr, nullptr, nullptr};
        int k = 0;
        for (const int* s : src) ptrs[k++] = s;
        poll_read_ints(c, ptrs, k);
        const int* got = reinterpret_cast<const int*>(c->h_poll);
        for (int i = 0; i < k; ++i) out[i] = got[i];
        return;
    }
    int* pin = reinterpret_cast<int*>(c->h_pin);
    size_t k = 0;
    for (const int* s : src) SSN_CUDA(cudaMemcpyAsync(pin + k++, s, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < k; ++i) out[i] = pin[i];
}

// ---- scans / sorts (CUB plumbing, sparse.cu) ----
void exclusive_scan_int(ssn_ctx* c, const int* in, int* out, int64_t n);   // out[n] = total if out has n+1: see impl
// out has n+1 entries: out[0]=0, out[i+1]=sum_{k<=i} in[k]; returns total (host, synchronises)
int64_t scan_counts_to_ptr(ssn_ctx* c, const int* counts, int* ptr, int64_t n);
// the same scan without the host read of the total (ptr[n] holds it on the device): for callers that know the
// total already, or that read several totals with one synchronisation (read_ints)
void scan_counts_async(ssn_ctx* c, const int* counts, int* ptr, int64_t n);
// stable sort of (key,value) int pairs by key, keys < key_limit
void stable_sort_pairs(ssn_ctx* c, const int* keys_in, int* keys_out, const int* vals_in,
                       int* vals_out, int64_t n, int key_limit);

// ---- device helpers ----
// 16-byte asynchronous global -> shared copies (LDGSTS): the staging buffers of the plan-wide kernels.  A thread waits for
// its OWN copies (cp_async_wait<N>: all but the N most recent groups) and reads back only what it copied itself.
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum, result valid in thread 0 (and broadcast through smem to all); blockDim <= 1024
__device__ __forceinline__ double block_sum(double v, double* smem32) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = (lane < nw) ? smem32[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) smem32[0] = t;
    }
    __syncthreads();
    t = smem32[0];
    return t;
}

}  // namespace ssn
