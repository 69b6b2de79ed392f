// sparse.cu -- CSR building blocks: scans / stable sorts (CUB plumbing), SpMV, deterministic
// transpose, fixed-order SpGEMM with shared-memory dense-window accumulators, sparse add,
// principal-submatrix extraction and deterministic reductions.
#include "sparse.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace ssn {

// ------------------------------------------------------------------ scans / sorts

namespace {
// out[i] = in[0] + ... + in[i-1] for i < n (and out[n] = total when with_total): ONE block, one launch, no
// temporary storage -- the AMG setup runs ~70 scans of a few thousand counts per hierarchy, where the two
// kernels + temporary allocation of cub::DeviceScan cost more than the scan itself.  In-place safe.
// Used up to ssn_ctx::small_scan_max counts (16384; env SSN_SMALL_SCAN_MAX); beyond that cub::DeviceScan.  (Round 1's form,
// every thread walking its own contiguous chunk of global memory, took 8 us at 16k counts: uncoalesced.)
// pub_host != null: the total also goes to mapped pinned host memory, followed by the sequence word the host polls
// (scan_counts_to_ptr: the read of the total rides in the scan instead of a publish kernel of its own).
__global__ void __launch_bounds__(1024) small_scan_kernel(const int* in, int* out, int n, int with_total, int* pub_host,
                                                          volatile unsigned long long* flag_host, unsigned long long seq) {
    // tiles of 4096 counts: coalesced load into shared memory, 4 consecutive counts per thread, warp + block scan, coalesced
    // store; the running total is carried from tile to tile (a tile's outputs overwrite only that tile's inputs: in-place safe)
    constexpr int kTile = 4096;
    __shared__ int tile[kTile];
    __shared__ int wsum[32];
    __shared__ int tile_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += kTile) {
        const int cnt = min(kTile, n - base);
        for (int i = threadIdx.x; i < kTile; i += 1024) tile[i] = (i < cnt) ? in[base + i] : 0;
        __syncthreads();
        const int i0 = 4 * (int)threadIdx.x;
        const int v0 = tile[i0], v1 = tile[i0 + 1], v2 = tile[i0 + 2], v3 = tile[i0 + 3];
        const int s = v0 + v1 + v2 + v3;
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = wsum[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
            wsum[lane] = wi - w;                             // exclusive prefix of the warp sums
            if (lane == 31) tile_total = wi;
        }
        __syncthreads();
        int run = carry + wsum[warp] + incl - s;
        tile[i0] = run; run += v0; tile[i0 + 1] = run; run += v1; tile[i0 + 2] = run; run += v2; tile[i0 + 3] = run;
        carry += tile_total;
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += 1024) out[base + i] = tile[i];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (with_total) out[n] = carry;
        if (pub_host != nullptr) { pub_host[0] = carry; __threadfence_system(); *flag_host = seq; }
    }
}
}  // namespace

// ------------------------------------------------------------------ polled device->host reads
namespace {
__global__ void __launch_bounds__(256) poll_publish_kernel(const unsigned* __restrict__ src, int nwords, unsigned* dst_host,
                                                           volatile unsigned long long* flag_host, unsigned long long seq) {
    for (int i = threadIdx.x; i < nwords; i += 256) dst_host[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *flag_host = seq;
}
__global__ void poll_publish_ints_kernel(const int* a, const int* b, const int* c2, const int* d, int k, int* dst_host,
                                         volatile unsigned long long* flag_host, unsigned long long seq) {
    if (threadIdx.x == 0) {
        if (k > 0) dst_host[0] = *a;
        if (k > 1) dst_host[1] = *b;
        if (k > 2) dst_host[2] = *c2;
        if (k > 3) dst_host[3] = *d;
        __threadfence_system();
        *flag_host = seq;
    }
}
void poll_wait(ssn_ctx* c, unsigned long long seq) {
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(c->h_poll + ssn_ctx::kPinDoubles);
    unsigned long spins = 0;
    while (*flag != seq) {
        if ((++spins & 0xffffu) == 0) {                    // a failed kernel upstream would never publish: ask the stream
            const cudaError_t e = cudaStreamQuery(c->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) throw Error(SSN_E_CUDA, std::string("device read: ") + cudaGetErrorString(e));
            if (e == cudaSuccess && *flag != seq) { SSN_CUDA(cudaStreamSynchronize(c->stream)); if (*flag != seq) throw Error(SSN_E_CUDA, "device read: the values never arrived"); }
        }
    }
}
}  // namespace

void poll_read(ssn_ctx* c, const void* dev, size_t bytes) {
    const unsigned long long seq = ++c->poll_seq;
    unsigned long long* dflag = reinterpret_cast<unsigned long long*>(c->d_poll + ssn_ctx::kPinDoubles);
    SSN_LAUNCH(c, poll_publish_kernel, 1, 256, 0, reinterpret_cast<const unsigned*>(dev), (int)(bytes / 4), reinterpret_cast<unsigned*>(c->d_poll), dflag, seq);
    c->launches--;                                         // plumbing, not a kernel of the path
    poll_wait(c, seq);
}
void poll_read_ints(ssn_ctx* c, const int* const* src, int k) {
    const unsigned long long seq = ++c->poll_seq;
    unsigned long long* dflag = reinterpret_cast<unsigned long long*>(c->d_poll + ssn_ctx::kPinDoubles);
    SSN_LAUNCH(c, poll_publish_ints_kernel, 1, 32, 0, src[0], src[1], src[2], src[3], k, reinterpret_cast<int*>(c->d_poll), dflag, seq);
    c->launches--;
    poll_wait(c, seq);
}

void scan_counts_async(ssn_ctx* c, const int* counts, int* ptr, int64_t n) {
    if (n == 0) { SSN_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int), c->stream)); return; }
    if (n <= c->small_scan_max) { SSN_LAUNCH(c, small_scan_kernel, 1, 1024, 0, counts, ptr, (int)n, 1, (int*)nullptr, (volatile unsigned long long*)nullptr, 0ull); return; }
    SSN_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int), c->stream));
    size_t tmp_bytes = 0;
    SSN_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, counts, ptr + 1, (int)n, c->stream));
    Buf<unsigned char> tmp(c, tmp_bytes);
    SSN_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, counts, ptr + 1, (int)n, c->stream));
    c->launches++;
}

int64_t scan_counts_to_ptr(ssn_ctx* c, const int* counts, int* ptr, int64_t n) {
    if (n > 0 && n <= c->small_scan_max && c->poll_reads && c->h_poll) {
        // one launch: the scan publishes its total to the host itself
        const unsigned long long seq = ++c->poll_seq;
        unsigned long long* dflag = reinterpret_cast<unsigned long long*>(c->d_poll + ssn_ctx::kPinDoubles);
        SSN_LAUNCH(c, small_scan_kernel, 1, 1024, 0, counts, ptr, (int)n, 1, reinterpret_cast<int*>(c->d_poll), (volatile unsigned long long*)dflag, seq);
        poll_wait(c, seq);
        return (int64_t)*reinterpret_cast<const int*>(c->h_poll);
    }
    scan_counts_async(c, counts, ptr, n);
    if (n == 0) return 0;
    return (int64_t)read_scalar(c, ptr + n);
}

void exclusive_scan_int(ssn_ctx* c, const int* in, int* out, int64_t n) {
    if (n == 0) return;
    if (n <= c->small_scan_max) { SSN_LAUNCH(c, small_scan_kernel, 1, 1024, 0, in, out, (int)n, 0, (int*)nullptr, (volatile unsigned long long*)nullptr, 0ull); return; }
    size_t tmp_bytes = 0;
    SSN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, c->stream));
    Buf<unsigned char> tmp(c, tmp_bytes);
    SSN_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, (int)n, c->stream));
    c->launches++;
}

void stable_sort_pairs(ssn_ctx* c, const int* keys_in, int* keys_out, const int* vals_in, int* vals_out,
                       int64_t n, int key_limit) {
    if (n == 0) return;
    int bits = 1;
    while (bits < 31 && (1ll << bits) < (long long)key_limit) ++bits;
    size_t tmp_bytes = 0;
    SSN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                             bits, c->stream));
    Buf<unsigned char> tmp(c, tmp_bytes);
    SSN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                             bits, c->stream));
    c->launches += 3;
}

// ------------------------------------------------------------------ small helpers

namespace {

__global__ void fill_double_kernel(double* p, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_int_kernel(int* p, int64_t n, int v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void iota_kernel(int* p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int)i;
}
inline int grid_for(int64_t n, int block = 256, int cap = 148 * 16) {
    int64_t g = (n + block - 1) / block; if (g < 1) g = 1; if (g > cap) g = cap; return (int)g;
}

constexpr int kRedBlocks = 296;
template <int MODE>   // 0 sum, 1 dot
__global__ void __launch_bounds__(256) reduce_stage1(const double* __restrict__ x, const double* __restrict__ y,
                                                     int64_t n, double* __restrict__ part) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += (MODE == 0) ? x[i] : x[i] * y[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) reduce_stage2(const double* __restrict__ part, int np, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) s += part[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}
__global__ void __launch_bounds__(256) count_u8_stage1(const uint8_t* __restrict__ x, int64_t n, double* __restrict__ part) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += x[i] ? 1.0 : 0.0;
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

}  // namespace

void fill_double(ssn_ctx* c, double* p, int64_t n, double v) { if (n > 0) SSN_LAUNCH(c, fill_double_kernel, grid_for(n), 256, 0, p, n, v); }
void fill_int(ssn_ctx* c, int* p, int64_t n, int v) { if (n > 0) SSN_LAUNCH(c, fill_int_kernel, grid_for(n), 256, 0, p, n, v); }
void iota_int(ssn_ctx* c, int* p, int64_t n) { if (n > 0) SSN_LAUNCH(c, iota_kernel, grid_for(n), 256, 0, p, n); }

double dev_sum(ssn_ctx* c, const double* x, int64_t n) {
    if (n <= 0) return 0.0;
    Buf<double> part(c, kRedBlocks + 1);
    const int g = grid_for(n, 256, kRedBlocks);
    SSN_LAUNCH(c, reduce_stage1<0>, g, 256, 0, x, nullptr, n, part.p);
    SSN_LAUNCH(c, reduce_stage2, 1, 256, 0, part.p, g, part.p + kRedBlocks);
    return read_scalar(c, part.p + kRedBlocks);
}
void dev_sum_async(ssn_ctx* c, const double* x, int64_t n, double* out_dev) {
    if (n <= 0) { SSN_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), c->stream)); return; }
    Buf<double> part(c, kRedBlocks + 1);
    const int g = grid_for(n, 256, kRedBlocks);
    SSN_LAUNCH(c, reduce_stage1<0>, g, 256, 0, x, nullptr, n, part.p);
    SSN_LAUNCH(c, reduce_stage2, 1, 256, 0, part.p, g, out_dev);
}
double dev_dot(ssn_ctx* c, const double* x, const double* y, int64_t n) {
    if (n <= 0) return 0.0;
    Buf<double> part(c, kRedBlocks + 1);
    const int g = grid_for(n, 256, kRedBlocks);
    SSN_LAUNCH(c, reduce_stage1<1>, g, 256, 0, x, y, n, part.p);
    SSN_LAUNCH(c, reduce_stage2, 1, 256, 0, part.p, g, part.p + kRedBlocks);
    return read_scalar(c, part.p + kRedBlocks);
}
int64_t dev_count_nonzero_u8(ssn_ctx* c, const uint8_t* x, int64_t n) {
    if (n <= 0) return 0;
    Buf<double> part(c, kRedBlocks + 1);
    const int g = grid_for(n, 256, kRedBlocks);
    SSN_LAUNCH(c, count_u8_stage1, g, 256, 0, x, n, part.p);
    SSN_LAUNCH(c, reduce_stage2, 1, 256, 0, part.p, g, part.p + kRedBlocks);
    return (int64_t)read_scalar(c, part.p + kRedBlocks);
}

// ------------------------------------------------------------------ SpMV

namespace {

template <int TPR, bool ADD>
__global__ void __launch_bounds__(256) spmv_kernel(int nrows, const int* __restrict__ ptr, const int* __restrict__ idx,
                                                   const double* __restrict__ val, const double* __restrict__ x,
                                                   double* __restrict__ y) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = gt / TPR, sub = gt % TPR;
    double s = 0.0;
    if (row < nrows) {
        const int e1 = ptr[row + 1];
        for (int e = ptr[row] + sub; e < e1; e += TPR) s = fma(val[e], x[idx[e]], s);
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < nrows && sub == 0) y[row] = ADD ? (y[row] + s) : s;
}

template <bool ADD>
void spmv_dispatch(ssn_ctx* c, const CsrView& A, const double* x, double* y) {
    if (A.nrows == 0) return;
    const double avg = (double)A.nnz / (double)A.nrows;
    const int n = A.nrows;
    if (avg <= 3.0)       SSN_LAUNCH(c, (spmv_kernel<2, ADD>), cdiv((int64_t)n * 2, 256), 256, 0, n, A.ptr, A.idx, A.val, x, y);
    else if (avg <= 6.0)  SSN_LAUNCH(c, (spmv_kernel<4, ADD>), cdiv((int64_t)n * 4, 256), 256, 0, n, A.ptr, A.idx, A.val, x, y);
    else if (avg <= 12.0) SSN_LAUNCH(c, (spmv_kernel<8, ADD>), cdiv((int64_t)n * 8, 256), 256, 0, n, A.ptr, A.idx, A.val, x, y);
    else if (avg <= 24.0) SSN_LAUNCH(c, (spmv_kernel<16, ADD>), cdiv((int64_t)n * 16, 256), 256, 0, n, A.ptr, A.idx, A.val, x, y);
    else                  SSN_LAUNCH(c, (spmv_kernel<32, ADD>), cdiv((int64_t)n * 32, 256), 256, 0, n, A.ptr, A.idx, A.val, x, y);
}

}  // namespace

void spmv(ssn_ctx* c, const CsrView& A, const double* x, double* y) { spmv_dispatch<false>(c, A, x, y); }
void spmv_add(ssn_ctx* c, const CsrView& A, const double* x, double* y) { spmv_dispatch<true>(c, A, x, y); }

// ------------------------------------------------------------------ structure helpers

namespace {

__global__ void expand_rows_kernel(int nrows, const int* __restrict__ ptr, int* __restrict__ rowidx) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= nrows) return;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) rowidx[e] = row;
}

__global__ void hist_kernel(const int* __restrict__ keys, int64_t n, int* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + keys[i], 1);
}

__global__ void transpose_gather_kernel(int64_t nnz, const int* __restrict__ perm, const int* __restrict__ rowidx,
                                        const double* __restrict__ val, int* __restrict__ oidx, double* __restrict__ oval) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = perm[i];
        oidx[i] = rowidx[e];
        oval[i] = val[e];
    }
}

__global__ void extract_diag_kernel(int nrows, const int* __restrict__ ptr, const int* __restrict__ idx,
                                    const double* __restrict__ val, double* __restrict__ diag) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    double d = 0.0;
    for (int e = ptr[row]; e < ptr[row + 1]; ++e) if (idx[e] == row) { d = val[e]; break; }
    diag[row] = d;
}

__global__ void count_nonzero_rows_kernel(int nrows, const int* __restrict__ ptr, const double* __restrict__ val,
                                          int* __restrict__ counts) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    int cnt = 0;
    for (int e = ptr[row]; e < ptr[row + 1]; ++e) cnt += (val[e] != 0.0);
    counts[row] = cnt;
}
__global__ void copy_nonzero_rows_kernel(int nrows, const int* __restrict__ ptr, const int* __restrict__ idx,
                                         const double* __restrict__ val, const int* __restrict__ optr,
                                         int* __restrict__ oidx, double* __restrict__ oval) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    int o = optr[row];
    for (int e = ptr[row]; e < ptr[row + 1]; ++e) if (val[e] != 0.0) { oidx[o] = idx[e]; oval[o] = val[e]; ++o; }
}

__global__ void extract_count_kernel(int nsel, const int* __restrict__ sel, const int* __restrict__ newidx,
                                     const int* __restrict__ ptr, const int* __restrict__ idx, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nsel) return;
    const int r = sel[i];
    int cnt = 0;
    for (int e = ptr[r] + lane; e < ptr[r + 1]; e += 32) cnt += (newidx[idx[e]] >= 0);
    cnt = warp_sum_int(cnt);
    if (lane == 0) counts[i] = cnt;
}
__global__ void extract_fill_kernel(int nsel, const int* __restrict__ sel, const int* __restrict__ newidx,
                                    const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                                    const int* __restrict__ optr, int* __restrict__ oidx, double* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nsel) return;
    const int r = sel[i];
    int o = optr[i];
    const int e0 = ptr[r], e1 = ptr[r + 1];
    for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        int ni = -1; double v = 0.0;
        if (e < e1) { ni = newidx[idx[e]]; v = val[e]; }
        const unsigned ball = __ballot_sync(0xffffffffu, ni >= 0);
        if (ni >= 0) {
            const int pos = o + __popc(ball & ((1u << lane) - 1u));
            oidx[pos] = ni; oval[pos] = v;
        }
        o += __popc(ball);
    }
}

// sparse add  C = A + alpha*B, one thread per row (rows of interpolation matrices are short)
template <bool FILL>
__global__ void sparse_add_kernel(int nrows, const int* __restrict__ ap, const int* __restrict__ ai, const double* __restrict__ av,
                                  double alpha, const int* __restrict__ bp, const int* __restrict__ bi, const double* __restrict__ bv,
                                  int* __restrict__ counts, const int* __restrict__ optr, int* __restrict__ oidx,
                                  double* __restrict__ oval) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    int ea = ap[row], ea1 = ap[row + 1], eb = bp[row], eb1 = bp[row + 1];
    int o = FILL ? optr[row] : 0, cnt = 0;
    while (ea < ea1 || eb < eb1) {
        const int ca = (ea < ea1) ? ai[ea] : 0x7fffffff;
        const int cb = (eb < eb1) ? bi[eb] : 0x7fffffff;
        double v; int col;
        if (ca == cb)      { v = __dadd_rn(av[ea], __dmul_rn(alpha, bv[eb])); col = ca; ++ea; ++eb; }
        else if (ca < cb)  { v = av[ea]; col = ca; ++ea; }
        else               { v = __dmul_rn(alpha, bv[eb]); col = cb; ++eb; }
        if (v != 0.0) {
            if (FILL) { oidx[o] = col; oval[o] = v; ++o; }
            ++cnt;
        }
    }
    if (!FILL) counts[row] = cnt;
}

}  // namespace

void expand_rows(ssn_ctx* c, const CsrView& A, int* rowidx) {
    if (A.nrows > 0 && A.nnz > 0) SSN_LAUNCH(c, expand_rows_kernel, cdiv((int64_t)A.nrows * 32, 256), 256, 0, A.nrows, A.ptr, rowidx);
}

void extract_diag(ssn_ctx* c, const CsrView& A, double* diag) {
    if (A.nrows > 0) SSN_LAUNCH(c, extract_diag_kernel, cdiv(A.nrows, 256), 256, 0, A.nrows, A.ptr, A.idx, A.val, diag);
}

Csr csr_alloc_from_counts(ssn_ctx* c, int nrows, int ncols, const int* counts) {
    Csr C; C.c = c; C.nrows = nrows; C.ncols = ncols;
    C.ptr.alloc(c, (size_t)nrows + 1);
    C.nnz = scan_counts_to_ptr(c, counts, C.ptr, nrows);
    C.idx.alloc(c, C.nnz); C.val.alloc(c, C.nnz);
    return C;
}

Csr csr_alloc_known(ssn_ctx* c, int nrows, int ncols, const int* counts, int64_t nnz) {
    Csr C; C.c = c; C.nrows = nrows; C.ncols = ncols;
    C.ptr.alloc(c, (size_t)nrows + 1);
    scan_counts_async(c, counts, C.ptr, nrows);
    C.nnz = nnz;
    C.idx.alloc(c, C.nnz); C.val.alloc(c, C.nnz);
    return C;
}

Csr csr_copy(ssn_ctx* c, const CsrView& A) {
    Csr C; C.c = c; C.nrows = A.nrows; C.ncols = A.ncols; C.nnz = A.nnz;
    C.ptr.alloc(c, (size_t)A.nrows + 1); C.idx.alloc(c, A.nnz); C.val.alloc(c, A.nnz);
    SSN_CUDA(cudaMemcpyAsync(C.ptr.p, A.ptr, sizeof(int) * ((size_t)A.nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
    if (A.nnz) {
        SSN_CUDA(cudaMemcpyAsync(C.idx.p, A.idx, sizeof(int) * A.nnz, cudaMemcpyDeviceToDevice, c->stream));
        SSN_CUDA(cudaMemcpyAsync(C.val.p, A.val, sizeof(double) * A.nnz, cudaMemcpyDeviceToDevice, c->stream));
    }
    return C;
}

Csr drop_zeros(ssn_ctx* c, const CsrView& A) {
    Buf<int> counts(c, A.nrows);
    if (A.nrows) SSN_LAUNCH(c, count_nonzero_rows_kernel, cdiv(A.nrows, 256), 256, 0, A.nrows, A.ptr, A.val, counts.p);
    Csr C = csr_alloc_from_counts(c, A.nrows, A.ncols, counts);
    if (A.nrows && C.nnz) SSN_LAUNCH(c, copy_nonzero_rows_kernel, cdiv(A.nrows, 256), 256, 0, A.nrows, A.ptr, A.idx, A.val, C.ptr.p, C.idx.p, C.val.p);
    return C;
}

Csr transpose(ssn_ctx* c, const CsrView& A) {
    Csr T; T.c = c; T.nrows = A.ncols; T.ncols = A.nrows; T.nnz = A.nnz;
    Buf<int> counts(c, (size_t)A.ncols); counts.zero();
    T.ptr.alloc(c, (size_t)A.ncols + 1);
    if (A.nnz > 0) SSN_LAUNCH(c, hist_kernel, grid_for(A.nnz), 256, 0, A.idx, A.nnz, counts.p);
    scan_counts_async(c, counts, T.ptr, A.ncols);                    // the total is A.nnz: no host read
    T.idx.alloc(c, A.nnz); T.val.alloc(c, A.nnz);
    if (A.nnz == 0) return T;
    Buf<int> rowidx(c, A.nnz), ent(c, A.nnz), keys_out(c, A.nnz), perm(c, A.nnz);
    expand_rows(c, A, rowidx);
    iota_int(c, ent, A.nnz);
    stable_sort_pairs(c, A.idx, keys_out, ent, perm, A.nnz, A.ncols > 1 ? A.ncols : 2);
    SSN_LAUNCH(c, transpose_gather_kernel, grid_for(A.nnz), 256, 0, A.nnz, perm.p, rowidx.p, A.val, T.idx.p, T.val.p);
    return T;
}

Csr sparse_add(ssn_ctx* c, const CsrView& A, double alpha, const CsrView& B) {
    SSN_REQUIRE(A.nrows == B.nrows && A.ncols == B.ncols, SSN_E_INVALID, "sparse_add: shape mismatch");
    Buf<int> counts(c, A.nrows);
    if (A.nrows) SSN_LAUNCH(c, sparse_add_kernel<false>, cdiv(A.nrows, 128), 128, 0, A.nrows, A.ptr, A.idx, A.val, alpha,
                            B.ptr, B.idx, B.val, counts.p, nullptr, nullptr, nullptr);
    Csr C = csr_alloc_from_counts(c, A.nrows, A.ncols, counts);
    if (A.nrows && C.nnz) SSN_LAUNCH(c, sparse_add_kernel<true>, cdiv(A.nrows, 128), 128, 0, A.nrows, A.ptr, A.idx, A.val, alpha,
                                     B.ptr, B.idx, B.val, nullptr, C.ptr.p, C.idx.p, C.val.p);
    return C;
}

Csr extract_principal(ssn_ctx* c, const CsrView& A, const int* sel, int nsel, const int* newidx) {
    Buf<int> counts(c, nsel);
    if (nsel) SSN_LAUNCH(c, extract_count_kernel, cdiv((int64_t)nsel * 32, 256), 256, 0, nsel, sel, newidx, A.ptr, A.idx, counts.p);
    Csr C = csr_alloc_from_counts(c, nsel, nsel, counts);
    if (nsel && C.nnz) SSN_LAUNCH(c, extract_fill_kernel, cdiv((int64_t)nsel * 32, 256), 256, 0, nsel, sel, newidx, A.ptr, A.idx,
                                  A.val, C.ptr.p, C.idx.p, C.val.p);
    return C;
}

// ------------------------------------------------------------------ SpGEMM

namespace {

constexpr int kMaxWindow = 8192;

__device__ __forceinline__ int lower_bound_dev(const int* __restrict__ a, int lo, int hi, int key) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// split[k*(nwin+1) + w] = first entry of B-row k with column >= w*W  (w = 0..nwin)
__global__ void spgemm_split_kernel(int nrowsB, const int* __restrict__ bp, const int* __restrict__ bi, int W, int nwin,
                                    int* __restrict__ split) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nrowsB * (nwin + 1)) return;
    const int k = (int)(t / (nwin + 1)), w = (int)(t % (nwin + 1));
    const int b0 = bp[k], b1 = bp[k + 1];
    split[t] = (w == 0) ? b0 : (w == nwin ? b1 : lower_bound_dev(bi, b0, b1, w * W));
}

constexpr int kSmallT = 256;          // rows with at most this many products (and A entries) take the warp path

// Upper bound of the number of entries of every segment (output row, column window), and the
// classification of the rows: a row with <= kSmallT products is SMALL (one warp, whole row in
// segment 0); the segments of the other rows are appended to big_list for the windowed kernel.
__global__ void spgemm_ub_kernel(int nrows, const int* __restrict__ ap, const int* __restrict__ ai,
                                 const int* __restrict__ bp, const int* __restrict__ split, int W, int nwin, int ncolsB,
                                 int small_ok, int* __restrict__ ub, int* __restrict__ big_list, int* __restrict__ nbig) {
    const int lane = threadIdx.x & 31;
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= (int64_t)nrows * nwin) return;
    const int row = (int)(item / nwin), win = (int)(item % nwin);
    long long s = 0, tot = 0;
    const int a0 = ap[row], a1 = ap[row + 1];
    for (int e = a0 + lane; e < a1; e += 32) {
        const int k = ai[e];
        const int full = bp[k + 1] - bp[k];
        tot += full;
        s += split ? (split[(size_t)k * (nwin + 1) + win + 1] - split[(size_t)k * (nwin + 1) + win]) : full;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
    if (lane != 0) return;
    if (small_ok && (a1 - a0) <= kSmallT && tot <= kSmallT) { ub[item] = (win == 0) ? (int)tot : 0; return; }
    const int wcols = (win == nwin - 1) ? (ncolsB - win * W) : W;
    ub[item] = (s < (long long)wcols) ? (int)s : wcols;
    big_list[atomicAdd(nbig, 1)] = (int)item;
}

// SMALL rows: one warp per output row.  The <= kSmallT products of the row are formed in
// Gustavson order t = 0,1,... (k ascending, then B's column order), rank-sorted by (column, t) in
// shared memory and summed sequentially inside every column segment -- the same additions in the
// same order as the dense-accumulator kernel, without its per-row latency chain.
__global__ void __launch_bounds__(128) spgemm_small_kernel(
    int nrows, const int* __restrict__ ap, const int* __restrict__ ai, const double* __restrict__ av,
    const int* __restrict__ bp, const int* __restrict__ bi, const double* __restrict__ bv, int nwin,
    const int* __restrict__ ubptr, int* __restrict__ tidx, double* __restrict__ tval, int* __restrict__ cnt_out,
    int* __restrict__ nbig_out) {
    // ubptr == nullptr: optimistic mode -- row r owns the fixed segment [r*kSmallT, (r+1)*kSmallT) and rows
    // that do not qualify are only counted in nbig_out (the caller then takes the general path)
    __shared__ int s_offs[4][kSmallT + 1];
    __shared__ int s_b0[4][kSmallT];
    __shared__ double s_a[4][kSmallT];
    __shared__ unsigned s_key[4][kSmallT];
    __shared__ double s_val[4][kSmallT];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int row = blockIdx.x * 4 + w; row < nrows; row += gridDim.x * 4) {
        const int a0 = ap[row], a1 = ap[row + 1], lenA = a1 - a0;
        if (lenA > kSmallT && nbig_out && lane == 0) atomicAdd(nbig_out, 1);
        if (lenA == 0 || lenA > kSmallT) continue;
        int running = 0;
        for (int base = 0; base < lenA; base += 32) {
            const int e = a0 + base + lane;
            int len = 0;
            if (e < a1) {
                const int k = ai[e];
                const int b0 = bp[k];
                len = bp[k + 1] - b0;
                s_b0[w][base + lane] = b0; s_a[w][base + lane] = av[e];
            }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (e < a1) s_offs[w][base + lane] = running + incl - len;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        const int T = running;
        if (T > kSmallT) { if (nbig_out && lane == 0) atomicAdd(nbig_out, 1); continue; }   // a big row: the windowed kernel owns it
        if (lane == 0) s_offs[w][lenA] = T;
        __syncwarp();
        for (int t = lane; t < T; t += 32) {
            int jl = 0, jh = lenA;                          // last j with offs[j] <= t
            while (jl < jh) { const int mid = (jl + jh) >> 1; if (s_offs[w][mid] <= t) jl = mid + 1; else jh = mid; }
            const int j = jl - 1;
            const int src = s_b0[w][j] + (t - s_offs[w][j]);
            s_key[w][t] = ((unsigned)bi[src] << 8) | (unsigned)t;
            s_val[w][t] = __dmul_rn(s_a[w][j], bv[src]);
        }
        // sort the products by (column, order of formation): the keys are distinct, so any sorting network gives THE order --
        // a bitonic network over the next power of two (log^2 steps; the rank sort it replaces was T^2/32 steps per lane)
        __syncwarp();                                       // the products of all lanes are in shared memory
        if (T <= 64) {
            // short rows (the common case): rank of each product among the row's products, two per lane held in registers
            unsigned k0 = 0u, k1 = 0u; double v0 = 0.0, v1 = 0.0; int r0 = 0, r1 = 0;
            const bool h0 = lane < T, h1 = lane + 32 < T;
            if (h0) { k0 = s_key[w][lane]; v0 = s_val[w][lane]; }
            if (h1) { k1 = s_key[w][lane + 32]; v1 = s_val[w][lane + 32]; }
            for (int u = 0; u < T; ++u) { const unsigned ku = s_key[w][u]; r0 += (ku < k0) ? 1 : 0; r1 += (ku < k1) ? 1 : 0; }
            __syncwarp();
            if (h0) { s_key[w][r0] = k0; s_val[w][r0] = v0; }
            if (h1) { s_key[w][r1] = k1; s_val[w][r1] = v1; }
            __syncwarp();
        } else {
        int Pw = 32;
        while (Pw < T) Pw <<= 1;
        for (int t = T + lane; t < Pw; t += 32) s_key[w][t] = 0xffffffffu;
        __syncwarp();
        for (int kk = 2; kk <= Pw; kk <<= 1) {
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                for (int i = lane; i < Pw; i += 32) {
                    const int ixj = i ^ jj;
                    if (ixj > i) {
                        const unsigned ka = s_key[w][i], kb = s_key[w][ixj];
                        const bool up = (i & kk) == 0;
                        if ((ka > kb) == up) {
                            s_key[w][i] = kb; s_key[w][ixj] = ka;
                            const double va = s_val[w][i]; s_val[w][i] = s_val[w][ixj]; s_val[w][ixj] = va;
                        }
                    }
                }
                __syncwarp();
            }
        }
        }
        const int out0 = ubptr ? ubptr[(size_t)row * nwin] : row * kSmallT;
        int written = 0;
        for (int base = 0; base < T; base += 32) {
            const int i = base + lane;
            bool keep = false; double acc = 0.0; int col = 0;
            if (i < T) {
                col = (int)(s_key[w][i] >> 8);
                const bool head = (i == 0) || ((int)(s_key[w][i - 1] >> 8) != col);
                if (head) {
                    acc = __dadd_rn(0.0, s_val[w][i]);
                    for (int j = i + 1; j < T && (int)(s_key[w][j] >> 8) == col; ++j) acc = __dadd_rn(acc, s_val[w][j]);
                    keep = (acc != 0.0);
                }
            }
            const unsigned ball = __ballot_sync(0xffffffffu, keep);
            if (keep) { const int pos = out0 + written + __popc(ball & ((1u << lane) - 1u)); tidx[pos] = col; tval[pos] = acc; }
            written += __popc(ball);
        }
        if (lane == 0) cnt_out[(size_t)row * nwin] = written;
        __syncwarp();
    }
}

// One block per work item = (output row, window of W output columns), handed out through an
// atomic counter.  Dense accumulator acc[W] in shared memory.  The A-row is walked in ascending k
// in chunks of THREADS entries; the B-rows of a chunk (restricted to the window, pre-multiplied
// by a_ik) are STAGED into shared memory by all threads at once -- the global-memory latency is
// paid once per chunk, not once per B-row -- and then accumulated group by group with one
// barrier between groups.  A group is one B-row, or a maximal run of single-entry B-rows with
// strictly increasing columns (the identity block of an interpolation matrix): inside a group
// all columns are distinct, so the adds are race-free, and consecutive groups are ordered by the
// barrier, which fixes the per-entry summation order (k ascending, multiply then add, no FMA).
template <int THREADS, int CAP>
__global__ void __launch_bounds__(THREADS, (THREADS >= 512 ? 2 : 4)) spgemm_numeric_kernel(
    int nrows, const int* __restrict__ ap, const int* __restrict__ ai, const double* __restrict__ av,
    const int* __restrict__ bp, const int* __restrict__ bi, const double* __restrict__ bv, const int* __restrict__ split,
    int ncolsB, int W, int nwin, const int* __restrict__ ubptr, int* __restrict__ tidx, double* __restrict__ tval,
    int* __restrict__ cnt_out, int* __restrict__ item_counter, const int* __restrict__ big_list,
    const int* __restrict__ nbig) {
    extern __shared__ unsigned char smem_raw[];
    double* acc = reinterpret_cast<double*>(smem_raw);
    double* s_prod = acc + W;
    int* s_col = reinterpret_cast<int*>(s_prod + CAP);
    constexpr int NW = THREADS / 32;
    constexpr int U = CAP / THREADS;
    static_assert(CAP % THREADS == 0, "CAP must be a multiple of THREADS");
    __shared__ int s_off[THREADS + 1];        // staged-entry offset of every B-row of the chunk
    __shared__ int s_lo[THREADS];             // first entry of the (window-restricted) B-row in bi/bv
    __shared__ int s_len[THREADS];
    __shared__ int s_c[THREADS];              // column of a single-entry row (else -1)
    __shared__ double s_a[THREADS];
    __shared__ int s_gstart[THREADS + 1];     // staged-entry offset of every group
    __shared__ int s_warp[NW], s_warp2[NW];
    __shared__ int s_total, s_groups, s_item;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nitems = *nbig;

    while (true) {
        if (tid == 0) { const int q = atomicAdd(item_counter, 1); s_item = (q < nitems) ? big_list[q] : -1; }
        __syncthreads();
        const int item = s_item;
        if (item < 0) break;
        const int row = item / nwin, win = item - row * nwin;
        const int w0 = win * W;
        const int wcols = (ncolsB - w0 < W) ? (ncolsB - w0) : W;
        const int a0 = ap[row], a1 = ap[row + 1];
        const int out0 = ubptr[item];
        int written = 0;
        if (a1 > a0) {
            for (int t = tid; t < wcols; t += THREADS) acc[t] = 0.0;
            __syncthreads();
            for (int eb = a0; eb < a1; eb += THREADS) {
                const int e = eb + tid;
                const int nb = (a1 - eb < THREADS) ? (a1 - eb) : THREADS;
                int lo = 0, len = 0, c1 = -1; double a = 0.0;
                if (e < a1) {
                    const int k = ai[e]; a = av[e];
                    int b0, b1;
                    if (split) { b0 = split[(size_t)k * (nwin + 1) + win]; b1 = split[(size_t)k * (nwin + 1) + win + 1]; }
                    else { b0 = bp[k]; b1 = bp[k + 1]; }
                    lo = b0; len = b1 - b0;
                    if (len == 1) c1 = bi[b0];
                }
                s_lo[tid] = lo; s_len[tid] = len; s_c[tid] = c1; s_a[tid] = a;
                __syncthreads();
                // a row opens a new group unless it extends a run of unit rows with increasing columns
                int ng = 0;
                if (len > 0) ng = (tid > 0 && len == 1 && s_len[tid - 1] == 1 && c1 > s_c[tid - 1]) ? 0 : 1;
                int il = len, ig = ng;                     // inclusive scans of len and ng
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int vl = __shfl_up_sync(0xffffffffu, il, o), vg = __shfl_up_sync(0xffffffffu, ig, o);
                    if (lane >= o) { il += vl; ig += vg; }
                }
                if (lane == 31) { s_warp[wid] = il; s_warp2[wid] = ig; }
                __syncthreads();
                if (wid == 0) {
                    int wl = (lane < NW) ? s_warp[lane] : 0, wg = (lane < NW) ? s_warp2[lane] : 0;
                    int xl = wl, xg = wg;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int vl = __shfl_up_sync(0xffffffffu, xl, o), vg = __shfl_up_sync(0xffffffffu, xg, o);
                        if (lane >= o) { xl += vl; xg += vg; }
                    }
                    if (lane < NW) { s_warp[lane] = xl - wl; s_warp2[lane] = xg - wg; }
                    if (lane == NW - 1) { s_total = xl; s_groups = xg; }
                }
                __syncthreads();
                const int off = s_warp[wid] + il - len;                 // exclusive
                const int gid = s_warp2[wid] + ig - ng;                 // groups before this row
                s_off[tid] = off;
                if (ng) s_gstart[gid] = off;
                const int total = s_total, groups = s_groups;
                if (tid == 0) { s_off[THREADS] = total; s_gstart[groups] = total; }
                __syncthreads();
                int g = 0;
                for (int P0 = 0; P0 < total; P0 += CAP) {
                    const int P1 = (P0 + CAP < total) ? (P0 + CAP) : total;
                    // ---- stage entries [P0, P1): entry t belongs to the last row j with s_off[j] <= t
                    int src[U]; double aj[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int t = P0 + u * THREADS + tid;
                        src[u] = -1; aj[u] = 0.0;
                        if (t < P1) {
                            int jl = 0, jh = nb;                                // upper_bound over s_off[0..nb)
                            while (jl < jh) { const int mid = (jl + jh) >> 1; if (s_off[mid] <= t) jl = mid + 1; else jh = mid; }
                            const int j = jl - 1;
                            src[u] = s_lo[j] + (t - s_off[j]); aj[u] = s_a[j];
                        }
                    }
                    int col[U]; double bval[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) { col[u] = 0; bval[u] = 0.0; if (src[u] >= 0) { col[u] = bi[src[u]]; bval[u] = bv[src[u]]; } }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (src[u] >= 0) { s_col[u * THREADS + tid] = col[u] - w0; s_prod[u * THREADS + tid] = __dmul_rn(aj[u], bval[u]); }
                    __syncthreads();
                    // ---- accumulate group by group (g persists across passes: a group may straddle two)
                    while (g < groups) {
                        int gs = s_gstart[g];
                        const int ge_full = s_gstart[g + 1];
                        if (gs >= P1) break;
                        if (gs < P0) gs = P0;
                        const int ge = (ge_full > P1) ? P1 : ge_full;
                        for (int t = gs + tid; t < ge; t += THREADS) {
                            const int cc = s_col[t - P0];
                            acc[cc] = __dadd_rn(acc[cc], s_prod[t - P0]);
                        }
                        __syncthreads();
                        if (ge_full > P1) break;
                        ++g;
                    }
                }
                __syncthreads();
            }
            // ---- compact the window in ascending column order (exact zeros dropped)
            const int per = (wcols + THREADS - 1) / THREADS;
            const int clo = (tid * per < wcols) ? tid * per : wcols, chi = (clo + per < wcols) ? (clo + per) : wcols;
            int cnt = 0;
            for (int t = clo; t < chi; ++t) if (acc[t] != 0.0) ++cnt;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            if (wid == 0) {
                int wv = (lane < NW) ? s_warp[lane] : 0;
                int wi = wv;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
                if (lane < NW) s_warp[lane] = wi - wv;
                if (lane == NW - 1) s_total = wi;
            }
            __syncthreads();
            int pos = out0 + s_warp[wid] + incl - cnt;
            for (int t = clo; t < chi; ++t)
                if (acc[t] != 0.0) { tidx[pos] = w0 + t; tval[pos] = acc[t]; ++pos; }
            written = s_total;
        }
        if (tid == 0) cnt_out[item] = written;
        __syncthreads();
    }
}

__global__ void sum_int64_kernel(const int* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += (unsigned long long)v[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// item segments (upper-bound layout) -> final CSR arrays; one warp per item
__global__ void compact_items_kernel(int64_t nitems, const int* __restrict__ ubptr, const int* __restrict__ cptr,
                                     const int* __restrict__ tidx, const double* __restrict__ tval,
                                     int* __restrict__ oidx, double* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= nitems) return;
    const int src = ubptr ? ubptr[item] : (int)item * kSmallT, dst = cptr[item], len = cptr[item + 1] - dst;
    for (int t = lane; t < len; t += 32) { oidx[dst + t] = tidx[src + t]; oval[dst + t] = tval[src + t]; }
}
__global__ void gather_rowptr_kernel(int nrows, int nwin, const int* __restrict__ cptr, int* __restrict__ optr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nrows) optr[i] = cptr[(size_t)i * nwin];
}

}  // namespace

namespace {
// 64-bit sum of the upper bounds of a row's work items
__global__ void row_ub64_kernel(int nrows, int nwin, const int* __restrict__ ub, long long* __restrict__ rowub) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    long long s = 0;
    for (int w = 0; w < nwin; ++w) s += ub[(size_t)i * nwin + w];
    rowub[i] = s;
}
__global__ void shift_ptr_kernel(int nrows, const int* __restrict__ src, int base, int* __restrict__ dst, int last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nrows) dst[i] = src[i] + base;
    if (last && i == nrows) dst[nrows] = src[nrows] + base;
}
}  // namespace

// C = A*B for a product whose intermediate upper bound exceeds ssn_ctx::spgemm_slab_limit: consecutive slabs of rows of A,
// each small enough for 32-bit offsets, multiplied on their own (same kernels, same per-row arithmetic) and concatenated.
static Csr spgemm_row_slabs(ssn_ctx* c, const CsrView& A, const CsrView& B, const int* ub, int nwin) {
    const int nrows = A.nrows;
    Buf<long long> rowub(c, nrows);
    SSN_LAUNCH(c, row_ub64_kernel, cdiv(nrows, 256), 256, 0, nrows, nwin, ub, rowub.p);
    std::vector<long long> h(nrows);
    SSN_CUDA(cudaMemcpyAsync(h.data(), rowub.p, sizeof(long long) * nrows, cudaMemcpyDeviceToHost, c->stream));
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    const int saved_site = c->spgemm_site;
    c->spgemm_site = -1;                                    // the slabs are not call sites of the hierarchy
    struct Restore { ssn_ctx* c; int s; ~Restore() { c->spgemm_site = s; } } restore{c, saved_site};
    std::vector<Csr> parts; std::vector<int> first;
    int64_t total = 0;
    for (int r0 = 0; r0 < nrows;) {
        int r1 = r0 + 1; long long acc = h[r0];
        while (r1 < nrows && acc + h[r1] < c->spgemm_slab_limit) acc += h[r1++];
        CsrView Av = A;
        Av.nrows = r1 - r0; Av.ptr = A.ptr + r0;
        Av.nnz = std::max<int64_t>(1, (int64_t)((double)A.nnz * (double)(r1 - r0) / (double)nrows));   // heuristics only
        parts.push_back(spgemm(c, Av, B));
        first.push_back(r0);
        total += parts.back().nnz;
        SSN_REQUIRE(total < ((int64_t)1 << 31), SSN_E_TOO_LARGE, "spgemm: the product itself exceeds the int32 index range");
        r0 = r1;
    }
    Csr C; C.c = c; C.nrows = nrows; C.ncols = B.ncols; C.nnz = total;
    C.ptr.alloc(c, (size_t)nrows + 1); C.idx.alloc(c, (size_t)total); C.val.alloc(c, (size_t)total);
    int64_t base = 0;
    for (size_t k = 0; k < parts.size(); ++k) {
        Csr& P = parts[k];
        const int rows = (int)P.nrows, last = (k + 1 == parts.size()) ? 1 : 0;
        SSN_LAUNCH(c, shift_ptr_kernel, cdiv(rows + 1, 256), 256, 0, rows, P.ptr.p, (int)base, C.ptr.p + first[k], last);
        if (P.nnz) {
            SSN_CUDA(cudaMemcpyAsync(C.idx.p + base, P.idx.p, sizeof(int) * P.nnz, cudaMemcpyDeviceToDevice, c->stream));
            SSN_CUDA(cudaMemcpyAsync(C.val.p + base, P.val.p, sizeof(double) * P.nnz, cudaMemcpyDeviceToDevice, c->stream));
        }
        base += P.nnz;
    }
    SSN_CUDA(cudaStreamSynchronize(c->stream));             // the parts are released on return
    return C;
}

Csr spgemm(ssn_ctx* c, const CsrView& A, const CsrView& B) {
    SSN_REQUIRE(A.ncols == B.nrows, SSN_E_INVALID, "spgemm: inner dimensions differ");
    const int nrows = A.nrows, ncolsB = B.ncols;
    Csr C; C.c = c; C.nrows = nrows; C.ncols = ncolsB;
    if (nrows == 0 || A.nnz == 0 || B.nnz == 0) {
        C.ptr.alloc(c, (size_t)nrows + 1); C.ptr.zero(); C.nnz = 0; C.idx.alloc(c, 0); C.val.alloc(c, 0);
        return C;
    }
    // ---- optimistic path (late SsN steps: every row is short): one warp per row writes into a fixed
    // kSmallT-entry segment, no upper-bound pass and a single host round trip; rows that do not
    // qualify are counted and, if there are any, the general path below redoes the product.
    const double avg_products = ((double)A.nnz / (double)nrows) * ((double)B.nnz / (double)(B.nrows > 0 ? B.nrows : 1));
    // Whether a row exceeds the warp path is only known afterwards; consecutive hierarchies of a solve repeat the same
    // products with nearly the same matrices, so the outcome of the optimistic attempt is remembered per call site
    // (ssn_ctx::spgemm_site: the n-th product since amg_setup started) and a site that failed last time goes straight to
    // the general path -- the attempt it skips costs a kernel over every row plus a host read.  Either path gives the
    // same product bit for bit.
    const int site = (c->spgemm_site >= 0 && c->spgemm_site < ssn_ctx::kSpgemmSites) ? c->spgemm_site++ : -1;
    const bool predicted_big = site >= 0 && c->spgemm_big[site] != 0;
    if (!predicted_big && ncolsB < (1 << 23) && (int64_t)nrows * kSmallT <= ((int64_t)1 << 24) && avg_products <= 192.0) {
        Buf<int> cnt(c, nrows), cptr(c, (size_t)nrows + 1), nbig(c, 1);
        Buf<int> tidx(c, (size_t)nrows * kSmallT); Buf<double> tval(c, (size_t)nrows * kSmallT);
        cnt.zero(); nbig.zero();
        int grid = cdiv(nrows, 4); if (grid > c->num_sms * 16) grid = c->num_sms * 16;
        SSN_LAUNCH(c, spgemm_small_kernel, grid, 128, 0, nrows, A.ptr, A.idx, A.val, B.ptr, B.idx, B.val, 1, nullptr, tidx.p, tval.p,
                   cnt.p, nbig.p);
        scan_counts_async(c, cnt, cptr, nrows);
        int h2[2];
        read_ints(c, {cptr.p + nrows, nbig.p}, h2);                       // the product's size and the fallback flag: one synchronisation
        const int64_t nnz = h2[0];
        if (site >= 0) c->spgemm_big[site] = (h2[1] != 0) ? 1 : 0;
        if (h2[1] == 0) {
            C.nnz = nnz;
            C.idx.alloc(c, C.nnz); C.val.alloc(c, C.nnz);
            if (C.nnz) SSN_LAUNCH(c, compact_items_kernel, cdiv((int64_t)nrows * 32, 256), 256, 0, (int64_t)nrows, nullptr, cptr.p, tidx.p,
                                  tval.p, C.idx.p, C.val.p);
            C.ptr = std::move(cptr);
            return C;
        }
    }
    // ---- work items: (row, window of W columns).  Few rows with a lot of work each are split into
    // narrow windows so that every SM gets items.
    int W = ((ncolsB + 31) / 32) * 32; if (W > kMaxWindow) W = kMaxWindow;
    int nwin = cdiv(ncolsB, W);
    const double work = (double)A.nnz * ((double)B.nnz / (double)(B.nrows > 0 ? B.nrows : 1));
    const int64_t want_items = 4 * (int64_t)c->num_sms;
    if ((int64_t)nrows * nwin < want_items && work > 2e6 && ncolsB > 512) {
        int nw = (int)cdiv(want_items, nrows);
        const int max_nw = cdiv(ncolsB, 256);
        if (nw > max_nw) nw = max_nw;
        W = ((cdiv(ncolsB, nw) + 31) / 32) * 32;
        nwin = cdiv(ncolsB, W);
    }
    SSN_REQUIRE((int64_t)nrows * nwin < ((int64_t)1 << 30), SSN_E_TOO_LARGE, "spgemm: too many work items");
    const int64_t nitems = (int64_t)nrows * nwin;
    Buf<int> split;
    if (nwin > 1) {
        split.alloc(c, (size_t)B.nrows * (nwin + 1));
        SSN_LAUNCH(c, spgemm_split_kernel, cdiv((int64_t)B.nrows * (nwin + 1), 256), 256, 0, B.nrows, B.ptr, B.idx, W, nwin, split.p);
    }
    const int* splitp = nwin > 1 ? split.p : nullptr;
    Buf<int> ub(c, nitems), ubptr(c, (size_t)nitems + 1), cnt(c, nitems), cptr(c, (size_t)nitems + 1);
    Buf<int> big_list(c, nitems), counters(c, 2);          // counters[0] = number of big items, [1] = work-item cursor
    counters.zero(); cnt.zero();
    const int small_ok = (ncolsB < (1 << 23)) ? 1 : 0;     // (column << 8 | t) must fit 32 bits
    SSN_LAUNCH(c, spgemm_ub_kernel, cdiv(nitems * 32, 256), 256, 0, nrows, A.ptr, A.idx, B.ptr, splitp, W, nwin, ncolsB, small_ok,
               ub.p, big_list.p, counters.p);
    if ((int64_t)nrows * (int64_t)ncolsB >= c->spgemm_slab_limit && nrows > 1) {   // the int32 scan could wrap: check in 64 bit
        Buf<unsigned long long> tot(c, 1); tot.zero();
        SSN_LAUNCH(c, sum_int64_kernel, 64, 256, 0, ub.p, nitems, tot.p);
        const unsigned long long t = read_scalar(c, tot.p);
        // more intermediate entries than 32-bit offsets (or the workspace) allow: the product is formed slab of rows by
        // slab of rows -- each slab a product of its own, bit for bit the rows of the whole product -- and concatenated
        if ((int64_t)t >= c->spgemm_slab_limit) return spgemm_row_slabs(c, A, B, ub.p, nwin);
    }
    const int64_t ub_total = scan_counts_to_ptr(c, ub, ubptr, nitems);
    Buf<int> tidx(c, (size_t)ub_total); Buf<double> tval(c, (size_t)ub_total);
    if (small_ok) {
        int grid = cdiv(nrows, 4); if (grid > c->num_sms * 16) grid = c->num_sms * 16;
        SSN_LAUNCH(c, spgemm_small_kernel, grid, 128, 0, nrows, A.ptr, A.idx, A.val, B.ptr, B.idx, B.val, nwin, ubptr.p, tidx.p, tval.p, cnt.p,
                   nullptr);
    }
    if (W > 2048) {
        constexpr int CAP = 2560;
        const size_t smem = (size_t)W * sizeof(double) + (size_t)CAP * (sizeof(double) + sizeof(int));
        auto kern = spgemm_numeric_kernel<512, CAP>;
        SSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (smem + 16 * 1024 <= 110 * 1024) ? 2 : 1;
        int64_t grid = (int64_t)c->num_sms * per_sm; if (grid > nitems) grid = nitems;
        SSN_LAUNCH(c, kern, (int)grid, 512, smem, nrows, A.ptr, A.idx, A.val, B.ptr, B.idx, B.val, splitp, ncolsB, W, nwin,
                   ubptr.p, tidx.p, tval.p, cnt.p, counters.p + 1, big_list.p, counters.p);
    } else {
        constexpr int CAP = 1024;
        const size_t smem = (size_t)W * sizeof(double) + (size_t)CAP * (sizeof(double) + sizeof(int));
        auto kern = spgemm_numeric_kernel<128, CAP>;
        SSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = (int)(200 * 1024 / (smem + 6 * 1024));
        per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
        int64_t grid = (int64_t)c->num_sms * per_sm; if (grid > nitems) grid = nitems;
        SSN_LAUNCH(c, kern, (int)grid, 128, smem, nrows, A.ptr, A.idx, A.val, B.ptr, B.idx, B.val, splitp, ncolsB, W, nwin,
                   ubptr.p, tidx.p, tval.p, cnt.p, counters.p + 1, big_list.p, counters.p);
    }
    C.nnz = scan_counts_to_ptr(c, cnt, cptr, nitems);
    C.idx.alloc(c, C.nnz); C.val.alloc(c, C.nnz);
    if (C.nnz) SSN_LAUNCH(c, compact_items_kernel, cdiv(nitems * 32, 256), 256, 0, nitems, ubptr.p, cptr.p, tidx.p, tval.p, C.idx.p, C.val.p);
    if (nwin == 1) C.ptr = std::move(cptr);
    else {
        C.ptr.alloc(c, (size_t)nrows + 1);
        SSN_LAUNCH(c, gather_rowptr_kernel, cdiv(nrows + 1, 256), 256, 0, nrows, nwin, cptr.p, C.ptr.p);
    }
    return C;
}

}  // namespace ssn
