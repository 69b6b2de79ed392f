// tests/emu/common.cuh -- HOST stand-in for csrc/common.cuh (test infrastructure only).
//
// There is no GPU in the build container.  Kernels that use nothing but thread indices, plain loads / stores,
// integer atomics, the *_rn arithmetic intrinsics and the warp votes can still be executed on the CPU: this header
// gives them the names they expect (blockIdx, threadIdx, __ballot_sync, ...), a Buf / Csr / CsrView with the same
// members as the real ones (backed by malloc), and an SSN_LAUNCH that runs every thread of the grid -- one after the
// other, or, when emu::warp_mode is set, the 32 lanes of each warp as 32 host threads that meet at the votes.
// A test copies the real .cu (and the real amg.cuh / sparse.cuh) next to this file and compiles it with g++, so the
// text that is checked is the text nvcc compiles.  It checks indexing, ordering and arithmetic; it does not check
// what only the hardware can (memory ordering between blocks, occupancy, cooperative launches).
#pragma once

#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "ssnamg.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__

typedef void* cudaStream_t;

namespace emu {
struct Dim { int x = 1, y = 1, z = 1; };
struct WarpShared { std::barrier<> bar{32}; unsigned votes[32]; };
inline bool warp_mode = false;                    // set by the harness before launching a kernel that votes
inline thread_local WarpShared* warp = nullptr;
}
inline thread_local emu::Dim blockIdx, threadIdx;
inline thread_local emu::Dim blockDim, gridDim;

// ---- arithmetic intrinsics: one IEEE operation each (compile with -ffp-contract=off)
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }     // launches are sequential or one warp at a time

// ---- warp votes (warp_mode only: the 32 lanes are 32 host threads)
inline unsigned __ballot_sync(unsigned, bool pred) {
    emu::WarpShared* w = emu::warp;
    w->votes[threadIdx.x & 31] = pred ? 1u : 0u;
    w->bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= w->votes[l] << l;
    w->bar.arrive_and_wait();
    return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }

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

struct ssn_ctx {
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    int num_sms = 148;
    bool device_setup = true;
};

namespace ssn {

template <class T>
struct Buf {
    ssn_ctx* c = nullptr; T* p = nullptr; size_t n = 0;
    Buf() = default;
    Buf(ssn_ctx* ctx, size_t count) { alloc(ctx, count); }
    Buf(const Buf&) = delete; Buf& operator=(const Buf&) = delete;
    Buf(Buf&& o) noexcept : c(o.c), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    Buf& operator=(Buf&& o) noexcept { if (this != &o) { reset(); c = o.c; p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~Buf() { reset(); }
    // filled with 0xA5 so that a read of something never written shows up as garbage, not as a lucky zero
    void alloc(ssn_ctx* ctx, size_t count) { reset(); c = ctx; n = count; p = (T*)std::malloc((count ? count : 1) * sizeof(T)); std::memset(p, 0xA5, (count ? count : 1) * sizeof(T)); }
    void reset() { if (p) { std::free(p); p = nullptr; n = 0; } }
    T* release() { T* r = p; p = nullptr; n = 0; return r; }
    void zero() { std::memset(p, 0, (n ? n : 1) * sizeof(T)); }
    operator T*() const { return p; }
    T* get() const { return p; }
};

struct Csr {
    ssn_ctx* c = nullptr;
    int64_t nrows = 0, ncols = 0, nnz = 0;
    Buf<int> ptr; Buf<int> idx; Buf<double> val;
    Csr() = default;
    Csr(Csr&&) = default; Csr& operator=(Csr&&) = default;
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

template <class K, class... A>
inline void emu_launch(ssn_ctx* c, K kernel, int grid, int block, A... args) {
    c->launches++;
    for (int b = 0; b < grid; ++b) {
        if (!emu::warp_mode) {
            for (int t = 0; t < block; ++t) {
                blockIdx.x = b; threadIdx.x = t; blockDim.x = block; gridDim.x = grid;
                kernel(args...);
            }
        } else {
            for (int w0 = 0; w0 < block; w0 += 32) {
                emu::WarpShared ws;
                std::vector<std::thread> lanes;
                for (int l = 0; l < 32; ++l)
                    lanes.emplace_back([&, l] {
                        emu::warp = &ws;
                        blockIdx.x = b; threadIdx.x = w0 + l; blockDim.x = block; gridDim.x = grid;
                        kernel(args...);
                        ws.bar.arrive_and_drop();          // a lane that has returned no longer takes part in votes
                    });
                for (auto& th : lanes) th.join();
            }
        }
    }
}
#define SSN_LAUNCH(ctx, kernel, grid, block, smem, ...) ::ssn::emu_launch((ctx), kernel, (grid), (block), __VA_ARGS__)

template <class T>
inline void read_back(ssn_ctx*, const T* dev, T* host, size_t count) { std::memcpy(host, dev, count * sizeof(T)); }
template <class T>
inline T read_scalar(ssn_ctx*, const T* dev) { return *dev; }

// the plumbing of sparse.cu that trifactor.cu calls, restated on the host (same contracts)
inline int64_t scan_counts_to_ptr(ssn_ctx*, const int* counts, int* ptr, int64_t n) {
    int64_t run = 0;
    for (int64_t i = 0; i < n; ++i) { const int v = counts[i]; ptr[i] = (int)run; run += v; }
    ptr[n] = (int)run;
    return run;
}
inline void exclusive_scan_int(ssn_ctx*, const int* in, int* out, int64_t n) { int run = 0; for (int64_t i = 0; i < n; ++i) { const int v = in[i]; out[i] = run; run += v; } }
inline void stable_sort_pairs(ssn_ctx*, const int* keys_in, int* keys_out, const int* vals_in, int* vals_out, int64_t n, int key_limit) {
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return keys_in[a] < keys_in[b]; });
    for (int64_t i = 0; i < n; ++i) {
        if (keys_in[order[i]] < 0 || keys_in[order[i]] >= key_limit) throw Error(SSN_E_INVALID, "stable_sort_pairs: key out of range");
        keys_out[i] = keys_in[order[i]]; vals_out[i] = vals_in[order[i]];
    }
}

}  // namespace ssn
