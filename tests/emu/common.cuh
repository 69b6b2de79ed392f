// tests/emu/common.cuh -- HOST stand-in for csrc/common.cuh (test infrastructure only).
//
// There is no GPU in the build container.  Kernels that use thread indices, plain loads / stores, shared memory,
// __syncthreads, warp votes / shuffles, atomics and the *_rn arithmetic intrinsics can still be executed on the CPU:
// this header gives them the names they expect, a Buf / Csr / CsrView with the same members as the real ones (backed
// by malloc), stubs of the few CUDA runtime calls the host code makes, and an SSN_LAUNCH that runs every thread of
// the grid: one after the other (emu::threaded = false: kernels without barriers, votes or shuffles), or block by
// block with one host thread per CUDA thread that meet at the barriers (emu::threaded = true).
// A test copies the real .cu / .cuh files next to this file and compiles them with g++, so the text that is checked
// is the text nvcc compiles (one mechanical edit: `extern __shared__ T x[];` becomes a pointer to emu::dyn_smem).
// It checks indexing, ordering and arithmetic; it does not check what only the hardware can (memory ordering between
// blocks, occupancy, cooperative launches, shared-memory capacity).
#pragma once
#ifndef SSN_EMU
#define SSN_EMU 1      // sources that need a host variant of a device-only construct (cluster special registers / barriers) test this
#endif

#include <algorithm>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "ssnamg.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static            /* blocks run one at a time, so one static instance is the block's shared memory */

#define __align__(n) alignas(n)

struct double2 { double x, y; };
struct int2 { int x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uchar2 { unsigned char x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline uchar2 make_uchar2(unsigned char x, unsigned char y) { return uchar2{x, y}; }
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum { cudaErrorNotReady = 600 };
inline cudaError_t cudaStreamQuery(cudaStream_t) { return 0; }
inline void __threadfence_system() {}
inline long long clock64() { return (long long)std::chrono::steady_clock::now().time_since_epoch().count(); }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

namespace emu {
struct Dim { int x = 1, y = 1, z = 1; };
struct WarpShared {
    std::barrier<> bar; unsigned long long slots[32];
    explicit WarpShared(int lanes) : bar(lanes) {}
};
struct BlockShared {
    std::barrier<> bar; int vote = 0;
    explicit BlockShared(int threads) : bar(threads) {}
};
inline bool threaded = false;                     // set by the harness: one host thread per CUDA thread, block by block
inline thread_local WarpShared* warp = nullptr;
inline thread_local BlockShared* block = nullptr;
alignas(64) inline unsigned char dyn_smem[256 * 1024];
// a thread-block cluster: all its blocks run at once, each with its own dynamic shared memory (cluster_smem[rank]) and
// block barrier, and meet at cluster_bar (emu_launch_cluster)
inline unsigned char* cluster_smem[16] = {nullptr};
inline std::barrier<>* cluster_bar = nullptr;
inline long host_reads = 0;                       // device->host reads (each one a stream synchronisation on the GPU)
inline void need_threads(const char* what) {
    if (!warp) { std::fprintf(stderr, "emu: %s needs emu::threaded = true\n", what); std::abort(); }
}
}
inline thread_local emu::Dim blockIdx, threadIdx;
inline thread_local emu::Dim blockDim, gridDim;

using std::max;
using std::min;
inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long max(long long a, int b) { return a > b ? a : b; }

// ---- arithmetic intrinsics: one IEEE operation each (compile with -ffp-contract=off)
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
// cp.async stand-ins: the copy happens at once
inline void cp_async16(void* smem, const void* g) { std::memcpy(smem, g, 16); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
inline int __double2hiint(double d) { long long b; std::memcpy(&b, &d, 8); return (int)(b >> 32); }
inline int __double2loint(double d) { long long b; std::memcpy(&b, &d, 8); return (int)(b & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) { const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &b, 8); return d; }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }

// ---- atomics (real ones: the threads of a block are host threads)
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double* p, double v) {
    unsigned long long* q = (unsigned long long*)p; unsigned long long o = __atomic_load_n(q, __ATOMIC_SEQ_CST), n;
    double od;
    do { std::memcpy(&od, &o, 8); const double nd = od + v; std::memcpy(&n, &nd, 8); } while (!__atomic_compare_exchange_n(q, &o, n, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return od;
}
inline int atomicMax(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
inline int atomicMin(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- block barrier, warp votes and shuffles (emu::threaded only)
inline void __syncthreads() { emu::need_threads("__syncthreads"); emu::block->bar.arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
    emu::need_threads("__syncthreads_or");
    emu::BlockShared* b = emu::block;
    if (pred) __atomic_store_n(&b->vote, 1, __ATOMIC_SEQ_CST);
    b->bar.arrive_and_wait();
    const int r = __atomic_load_n(&b->vote, __ATOMIC_SEQ_CST);
    b->bar.arrive_and_wait();
    if (threadIdx.x == 0) __atomic_store_n(&b->vote, 0, __ATOMIC_SEQ_CST);
    b->bar.arrive_and_wait();
    return r;
}
inline int __syncthreads_and(int pred) { return !__syncthreads_or(!pred); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::need_threads("__syncwarp"); emu::warp->bar.arrive_and_wait(); emu::warp->bar.arrive_and_wait(); }
template <class T> inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    emu::need_threads("a warp shuffle / vote");
    emu::WarpShared* w = emu::warp;
    unsigned long long bits = 0; std::memcpy(&bits, &v, sizeof(T));
    w->slots[threadIdx.x & 31] = bits;
    w->bar.arrive_and_wait();
    T r; std::memcpy(&r, &w->slots[src_lane & 31], sizeof(T));
    w->bar.arrive_and_wait();
    return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_exchange(v, (threadIdx.x & 31) ^ o); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned o) { const int l = threadIdx.x & 31; return emu_exchange(v, l >= (int)o ? l - (int)o : l); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned o) { const int l = threadIdx.x & 31; return emu_exchange(v, l + (int)o < 32 ? l + (int)o : l); }
inline unsigned __ballot_sync(unsigned, bool pred) {
    emu::need_threads("__ballot_sync");
    emu::WarpShared* w = emu::warp;
    w->slots[threadIdx.x & 31] = pred ? 1ull : 0ull;
    w->bar.arrive_and_wait();
    unsigned m = 0;
    const int lanes = std::min(32, blockDim.x - (threadIdx.x & ~31));
    for (int l = 0; l < lanes; ++l) m |= (unsigned)w->slots[l] << l;
    w->bar.arrive_and_wait();
    return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
inline bool __all_sync(unsigned mask, bool pred) { return __ballot_sync(mask, !pred) == 0u; }

namespace ssn {

struct Error : std::exception {
    int code; std::string msg;
    Error(int c, std::string m) : code(c), msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
#define SSN_REQUIRE(cond, code, text) do { if (!(cond)) throw ::ssn::Error((code), (text)); } while (0)
#define SSN_CUDA(call) do { (void)(call); } while (0)

struct Hierarchy;
}  // namespace ssn

// the fields of the real ssn_ctx that the emulated sources read
struct ssn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int num_sms = 2;                      // small grids: every block costs host threads
    size_t smem_optin = 227 * 1024;
    uint32_t* mt_state = nullptr;
    int64_t rng_drawn = 0;
    ssn::Hierarchy* hier = nullptr;
    bool no_cluster = true;
    int64_t persist_max_nnz = (int64_t)1 << 40;
    bool persist = true, dense_tail = true;
    int plan_waves = 0;                   // SSN_PLAN_WAVES: waves of blocks of the plan-wide reduction kernels (0: two)
    bool tg_cluster = true;               // SSN_TG_CLUSTER=0: twogrid_bigph's iteration loop kernel by kernel (coarse PCG: the grid-wide pcg_kernel)
    bool plan_stage = true;
    int ls_max_nt = 128; bool ls_screen = true; double ls_last_density = -1.0; int ls_last_ll = -1;
    int small_scan_max = 1 << 14;
    bool device_setup = true;
    static constexpr int kSpgemmSites = 64;
    int last_dsm_halo = -1;               // halo entries of CTA 0 in the last dsm_solve_kernel launch (diagnostic)
    int64_t spgemm_slab_limit = (int64_t)1 << 30;   // sparse products with more intermediate entries are formed in slabs of rows
    unsigned spgemm_epoch = 0;
    int spgemm_site = -1;                 // >= 0 inside amg_setup: index of the next sparse product of this hierarchy
    unsigned char spgemm_big[kSpgemmSites] = {0};   // 1: the optimistic warp-path attempt of that product failed last time
    bool stage_dense = true;              // SSN_STAGE_DENSE=0: the grid-wide solve kernel gathers from L2 on dense levels too
    bool mis_cluster = true;              // SSN_MIS_CLUSTER=0: the MIS rounds of mis_set.m launch by launch (one host read per round)
    bool fused_setup = true, cluster_solve = true;     // the emulated hierarchies go through the fused kernel (opt-in in the library)
    int64_t cluster_max_nnz = (int64_t)1 << 20;
    int dense_max_n = 2048;
    bool ktimer = false, prof = false;
    double* h_pin = nullptr;
    static constexpr int kPinDoubles = 4096;
    bool poll_reads = false; double* h_poll = nullptr; double* d_poll = nullptr; unsigned long long poll_seq = 0;   // polled reads: GPU only
    ssn_ctx() { mt_state = (uint32_t*)std::calloc(625, sizeof(uint32_t)); h_pin = (double*)std::calloc(kPinDoubles, sizeof(double)); }
    ~ssn_ctx() { std::free(mt_state); std::free(h_pin); }
    ssn_ctx(const ssn_ctx&) = delete; ssn_ctx& operator=(const ssn_ctx&) = delete;
};

namespace ssn {

template <class T>
struct Buf {
    ssn_ctx* c = nullptr; T* p = nullptr; size_t n = 0;
    bool owned = true;                    // false: a view (Buf::view), never freed here
    Buf() = default;
    Buf(ssn_ctx* ctx, size_t count) { alloc(ctx, count); }
    Buf(const Buf&) = delete; Buf& operator=(const Buf&) = delete;
    Buf(Buf&& o) noexcept : c(o.c), p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; }
    Buf& operator=(Buf&& o) noexcept { if (this != &o) { reset(); c = o.c; p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; } return *this; }
    ~Buf() { reset(); }
    static Buf view(ssn_ctx* ctx, T* ptr, size_t count) { Buf b; b.c = ctx; b.p = ptr; b.n = count; b.owned = false; return b; }
    // filled with 0xA5 so that a read of something never written shows up as garbage, not as a lucky zero
    void alloc(ssn_ctx* ctx, size_t count) { reset(); c = ctx; n = count; owned = true; p = (T*)std::malloc((count ? count : 1) * sizeof(T)); std::memset((void*)p, 0xA5, (count ? count : 1) * sizeof(T)); }
    void reset() { if (p && owned) std::free((void*)p); p = nullptr; n = 0; owned = true; }
    T* release() { T* r = p; p = nullptr; n = 0; return r; }
    void zero() { std::memset((void*)p, 0, (n ? n : 1) * sizeof(T)); }
    operator T*() const { return p; }
    T* get() const { return p; }
};

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
};

struct CsrView {
    int nrows = 0, ncols = 0; int64_t nnz = 0;
    const int* ptr = nullptr; const int* idx = nullptr; const double* val = nullptr;
    CsrView() = default;
    CsrView(const ssn_csr& a) : nrows((int)a.nrows), ncols((int)a.ncols), nnz(a.nnz),
        ptr((const int*)a.rowptr_dev), idx((const int*)a.colidx_dev), val((const double*)a.val_dev) {}
    CsrView(const Csr& a) : nrows((int)a.nrows), ncols((int)a.ncols), nnz(a.nnz), ptr(a.ptr.p), idx(a.idx.p), val(a.val.p) {}
};

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// the library's phase profiler, counting instead of timing: launches and device->host reads per named phase
inline std::map<std::string, std::pair<long, long>> phase_counts;
struct Phase {
    ssn_ctx* c; const char* name; long l0, r0;
    Phase(ssn_ctx* ctx, const char* n) : c(ctx), name(n), l0((long)ctx->launches), r0(emu::host_reads) {}
    ~Phase() { auto& a = phase_counts[name]; a.first += (long)c->launches - l0; a.second += emu::host_reads - r0; }
};
struct KernelTimer { explicit KernelTimer(ssn_ctx*) {} };

template <class K, class... A>
inline void emu_launch(ssn_ctx* c, K kernel, dim3 grid, int block, A... args) {
    c->launches++;
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned b = 0; b < grid.x; ++b) {
        auto set_ids = [&](int t) {
            blockIdx.x = (int)b; blockIdx.y = (int)by; threadIdx.x = t; blockDim.x = block; gridDim.x = (int)grid.x; gridDim.y = (int)grid.y;
        };
        if (!emu::threaded) {
            for (int t = 0; t < block; ++t) { set_ids(t); kernel(args...); }
        } else {
            emu::BlockShared bs(block);
            std::vector<std::unique_ptr<emu::WarpShared>> ws;
            for (int w0 = 0; w0 < block; w0 += 32) ws.emplace_back(new emu::WarpShared(std::min(32, block - w0)));
            std::vector<std::thread> threads;
            threads.reserve((size_t)block);
            for (int t = 0; t < block; ++t)
                threads.emplace_back([&, t] {
                    emu::block = &bs; emu::warp = ws[(size_t)t / 32].get();
                    set_ids(t);
                    kernel(args...);
                    emu::warp->bar.arrive_and_drop();       // a thread that has returned no longer takes part in barriers
                    bs.bar.arrive_and_drop();
                });
            for (auto& th : threads) th.join();
        }
    }
}
// one cluster of `ncta` blocks, all resident at once: ncta * block host threads; dynamic shared memory of `smem` bytes per
// block at emu::cluster_smem[blockIdx.x] (the kernel must not use static __shared__ variables: they would be shared)
template <class K, class... A>
inline void emu_launch_cluster(ssn_ctx* c, K kernel, int ncta, int block, size_t smem, A... args) {
    c->launches++;
    constexpr size_t kCanary = 4096;                      // untouched bytes behind every block's shared memory, checked after the kernel
    std::vector<std::vector<unsigned char>> mem((size_t)ncta, std::vector<unsigned char>(smem + 64 + kCanary, (unsigned char)0xA5));
    std::barrier<> cbar(ncta * block);
    emu::cluster_bar = &cbar;
    std::vector<std::unique_ptr<emu::BlockShared>> bs;
    std::vector<std::unique_ptr<emu::WarpShared>> ws;
    const int wpb = (block + 31) / 32;
    for (int b = 0; b < ncta; ++b) {
        emu::cluster_smem[b] = mem[(size_t)b].data() + ((64 - ((uintptr_t)mem[(size_t)b].data() & 63)) & 63);
        bs.emplace_back(new emu::BlockShared(block));
        for (int w0 = 0; w0 < block; w0 += 32) ws.emplace_back(new emu::WarpShared(std::min(32, block - w0)));
    }
    std::vector<std::thread> threads;
    threads.reserve((size_t)ncta * block);
    for (int b = 0; b < ncta; ++b)
        for (int t = 0; t < block; ++t)
            threads.emplace_back([&, b, t] {
                emu::block = bs[(size_t)b].get(); emu::warp = ws[(size_t)b * wpb + t / 32].get();
                blockIdx.x = b; blockIdx.y = 0; threadIdx.x = t; blockDim.x = block; gridDim.x = ncta; gridDim.y = 1;
                kernel(args...);
                emu::warp->bar.arrive_and_drop();
                emu::block->bar.arrive_and_drop();
                cbar.arrive_and_drop();
            });
    for (auto& th : threads) th.join();
    for (int b = 0; b < ncta; ++b)
        for (size_t i = 0; i < kCanary; ++i)
            if (emu::cluster_smem[b][smem + i] != (unsigned char)0xA5) {
                std::fprintf(stderr, "emu: block %d wrote %zu bytes past its %zu bytes of shared memory\n", b, i, smem);
                std::abort();
            }
    emu::cluster_bar = nullptr;
}
#define SSN_LAUNCH(ctx, kernel, grid, block, smem, ...) ::ssn::emu_launch((ctx), kernel, dim3(grid), (int)(block), __VA_ARGS__)

inline void upload_small(ssn_ctx*, void* dst_dev, const void* src_host, size_t bytes) { std::memcpy(dst_dev, src_host, bytes); }
template <class T>
inline void read_back(ssn_ctx*, const T* dev, T* host, size_t count) { ++emu::host_reads; std::memcpy(host, dev, count * sizeof(T)); }
template <class T>
inline T read_scalar(ssn_ctx*, const T* dev) { ++emu::host_reads; return *dev; }
inline void read_ints(ssn_ctx*, std::initializer_list<const int*> src, int* out) { ++emu::host_reads; size_t k = 0; for (const int* s : src) out[k++] = *s; }

// scans / sorts: declared here like in the real common.cuh.  Defined by the emulated sparse.cu when a harness compiles
// it (over tests/emu/cub).
void exclusive_scan_int(ssn_ctx* c, const int* in, int* out, int64_t n);
int64_t scan_counts_to_ptr(ssn_ctx* c, const int* counts, int* ptr, int64_t n);
void scan_counts_async(ssn_ctx* c, const int* counts, int* ptr, int64_t n);
void stable_sort_pairs(ssn_ctx* c, const int* keys_in, int* keys_out, const int* vals_in, int* vals_out, int64_t n, int key_limit);

inline double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
inline int warp_sum_int(int v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
inline double block_sum(double v, double* smem32) {
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
