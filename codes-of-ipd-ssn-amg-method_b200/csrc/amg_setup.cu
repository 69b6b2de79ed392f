// amg_setup.cu -- AMG setup on the device: MATLAB random stream (MT19937), strength of
// connection, the iFEM-style randomised MIS C/F split, the greedy cf_split, interpolation and
// Galerkin coarse operators, and the Class_AMG hierarchy.  Everything that decides a sparsity
// pattern or a C/F split uses explicitly rounded arithmetic (__dmul_rn/__dadd_rn/__ddiv_rn) in
// the oracle's order, so those decisions are bit-exact.
#include "amg.cuh"

namespace ssn {

// =================================================================== MT19937 (K17)

namespace {

__global__ void mt_seed_kernel(uint32_t* st, uint32_t seed) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t prev = seed;
    st[0] = prev;
    for (int i = 1; i < 624; ++i) { prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i; st[i] = prev; }
    st[624] = 624u;
}

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// genrand_res53 stream; the 624-word state regenerates in three data-parallel phases.
__global__ void __launch_bounds__(256) mt_rand_kernel(uint32_t* __restrict__ st, long long count, double* out) {
    __shared__ uint32_t mt[624];
    const int tid = threadIdx.x;
    for (int i = tid; i < 624; i += 256) mt[i] = st[i];
    int mti = (int)st[624];
    __syncthreads();
    uint32_t* words = reinterpret_cast<uint32_t*>(out);
    const long long total = 2 * count;
    long long produced = 0;
    while (produced < total) {
        if (mti >= 624) {
            uint32_t v = 0;
            if (tid < 227) v = mt_twist(mt[tid], mt[tid + 1], mt[tid + 397]);
            __syncthreads();
            if (tid < 227) mt[tid] = v;
            __syncthreads();
            if (tid < 227) v = mt_twist(mt[227 + tid], mt[228 + tid], mt[tid]);
            __syncthreads();
            if (tid < 227) mt[227 + tid] = v;
            __syncthreads();
            if (tid < 170) v = mt_twist(mt[454 + tid], mt[(455 + tid) % 624], mt[227 + tid]);
            __syncthreads();
            if (tid < 170) mt[454 + tid] = v;
            __syncthreads();
            mti = 0;
        }
        const int avail = 624 - mti;
        const long long rem = total - produced;
        const int take = rem < (long long)avail ? (int)rem : avail;
        for (int t = tid; t < take; t += 256) {
            uint32_t y = mt[mti + t];
            y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
            words[produced + t] = y;
        }
        mti += take; produced += take;
        __syncthreads();
    }
    for (int i = tid; i < 624; i += 256) st[i] = mt[i];
    if (tid == 0) st[624] = (uint32_t)mti;
    __syncthreads();
    for (long long i = tid; i < count; i += 256) {
        const uint32_t a = words[2 * i] >> 5, b = words[2 * i + 1] >> 6;
        const double d = ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
        out[i] = d;
    }
}

}  // namespace

void rng_reset(ssn_ctx* c, uint32_t seed) {
    SSN_LAUNCH(c, mt_seed_kernel, 1, 32, 0, c->mt_state, seed);
    c->rng_drawn = 0;
}

void rng_rand(ssn_ctx* c, int64_t count, double* out_dev) {
    if (count <= 0) return;
    SSN_LAUNCH(c, mt_rand_kernel, 1, 256, 0, c->mt_state, (long long)count, out_dev);
    c->rng_drawn += count;
}

// =================================================================== strength (K7)

namespace {

// max_row(i) = max_j A0(i,j) with A0 = D - A, implicit zeros included; <= 0 -> inf (strength.m:9-10)
__global__ void maxrow_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                              const double* __restrict__ val, double* __restrict__ maxrow) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    double mx = 0.0;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32)
        if (idx[e] != row) mx = fmax(mx, -val[e]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) maxrow[row] = (mx <= 0.0) ? INFINITY : mx;
}

__device__ __forceinline__ double strength_value(int i, int j, double a, const double* __restrict__ maxrow, int which) {
    const double s0 = -a;                                   // D - A off the diagonal
    const double den = (which == 1) ? maxrow[i] : fmin(maxrow[i], maxrow[j]);
    return __ddiv_rn(s0, den);
}

// flags[e] = 1 iff entry e is an off-diagonal nonzero whose strength value is >= theta.
// Also: deg[j] += 1 for flagged (i,j) (column counts, mis_set.m:28) and rowcnt[i] (mis_set.m:67).
__global__ void strength_flags_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                      const double* __restrict__ val, const double* __restrict__ maxrow, double theta,
                                      int which, uint8_t* __restrict__ flags, int* __restrict__ deg,
                                      int* __restrict__ rowcnt) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    int cnt = 0;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) {
        const int j = idx[e];
        const double a = val[e];
        uint8_t f = 0;
        if (j != row && a != 0.0) {
            const double sa = strength_value(row, j, a, maxrow, which);
            f = (sa >= theta) ? 1 : 0;
        }
        flags[e] = f;
        if (f) { ++cnt; if (deg) atomicAdd(deg + j, 1); }
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0 && rowcnt) rowcnt[row] = cnt;
}

__global__ void strength_count_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                      const double* __restrict__ val, const double* __restrict__ maxrow, int which,
                                      int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    int cnt = 0;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) {
        const int j = idx[e]; const double a = val[e];
        if (j != row && a != 0.0 && strength_value(row, j, a, maxrow, which) != 0.0) ++cnt;
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0) counts[row] = cnt;
}
__global__ void strength_fill_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                     const double* __restrict__ val, const double* __restrict__ maxrow, int which,
                                     const int* __restrict__ optr, int* __restrict__ oidx, double* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    int o = optr[row];
    const int e0 = ptr[row], e1 = ptr[row + 1];
    for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        bool keep = false; double sa = 0.0; int j = 0;
        if (e < e1) {
            j = idx[e]; const double a = val[e];
            if (j != row && a != 0.0) { sa = strength_value(row, j, a, maxrow, which); keep = (sa != 0.0); }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        if (keep) { const int pos = o + __popc(ball & ((1u << lane) - 1u)); oidx[pos] = j; oval[pos] = sa; }
        o += __popc(ball);
    }
}

__global__ void flags_count_kernel(int n, const int* __restrict__ ptr, const uint8_t* __restrict__ flags, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    int cnt = 0;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) cnt += flags[e];
    cnt = warp_sum_int(cnt);
    if (lane == 0) counts[row] = cnt;
}
__global__ void flags_fill_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                  const uint8_t* __restrict__ flags, const int* __restrict__ optr,
                                  int* __restrict__ oidx, double* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    int o = optr[row];
    const int e0 = ptr[row], e1 = ptr[row + 1];
    for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        const bool keep = (e < e1) && flags[e];
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        if (keep) { const int pos = o + __popc(ball & ((1u << lane) - 1u)); oidx[pos] = idx[e]; oval[pos] = 1.0; }
        o += __popc(ball);
    }
}

}  // namespace

static void compute_maxrow(ssn_ctx* c, const CsrView& A, double* maxrow) {
    if (A.nrows) SSN_LAUNCH(c, maxrow_kernel, cdiv((int64_t)A.nrows * 32, 256), 256, 0, A.nrows, A.ptr, A.idx, A.val, maxrow);
}

Csr strength_matrix(ssn_ctx* c, const CsrView& A, int which) {
    SSN_REQUIRE(A.nrows == A.ncols, SSN_E_NOT_SQUARE, "strength: matrix must be square");
    const int n = A.nrows;
    Buf<double> maxrow(c, n); Buf<int> counts(c, n);
    compute_maxrow(c, A, maxrow);
    if (n) SSN_LAUNCH(c, strength_count_kernel, cdiv((int64_t)n * 32, 256), 256, 0, n, A.ptr, A.idx, A.val, maxrow.p, which, counts.p);
    Csr S = csr_alloc_from_counts(c, n, n, counts);
    if (n && S.nnz) SSN_LAUNCH(c, strength_fill_kernel, cdiv((int64_t)n * 32, 256), 256, 0, n, A.ptr, A.idx, A.val, maxrow.p, which,
                               S.ptr.p, S.idx.p, S.val.p);
    return S;
}

static void strength_flags_full(ssn_ctx* c, const CsrView& A, double theta, uint8_t* flags, int* deg, int* rowcnt) {
    const int n = A.nrows;
    Buf<double> maxrow(c, n);
    compute_maxrow(c, A, maxrow);
    if (n) SSN_LAUNCH(c, strength_flags_kernel, cdiv((int64_t)n * 32, 256), 256, 0, n, A.ptr, A.idx, A.val, maxrow.p, theta, 2,
                      flags, deg, rowcnt);
}

void strength_flags(ssn_ctx* c, const CsrView& A, double theta, uint8_t* as_flags) {
    strength_flags_full(c, A, theta, as_flags, nullptr, nullptr);
}

Csr flags_to_csr(ssn_ctx* c, const CsrView& A, const uint8_t* flags) {
    const int n = A.nrows;
    Buf<int> counts(c, n);
    if (n) SSN_LAUNCH(c, flags_count_kernel, cdiv((int64_t)n * 32, 256), 256, 0, n, A.ptr, flags, counts.p);
    Csr S = csr_alloc_from_counts(c, n, A.ncols, counts);
    if (n && S.nnz) SSN_LAUNCH(c, flags_fill_kernel, cdiv((int64_t)n * 32, 256), 256, 0, n, A.ptr, A.idx, flags, S.ptr.p, S.idx.p, S.val.p);
    return S;
}

// =================================================================== MIS C/F split (K8)

namespace {

__global__ void mis_init_kernel(int n, const int* __restrict__ deg, const int* __restrict__ rank,
                                const double* __restrict__ rnd, double* __restrict__ degf,
                                uint8_t* __restrict__ isC, uint8_t* __restrict__ isF) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = deg[i];
    // deg(idx) = deg(idx) + 0.1*rand(sum(idx),1), draws in ascending node order (mis_set.m:35)
    degf[i] = (d > 0) ? __dadd_rn((double)d, __dmul_rn(0.1, rnd[rank[i]])) : 0.0;
    isC[i] = 0;
    isF[i] = (d == 0) ? 1 : 0;                               // mis_set.m:40
}
__global__ void positive_flag_kernel(int n, const int* __restrict__ deg, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = deg[i] > 0 ? 1 : 0;
}
__global__ void mis_mark_kernel(int n, const double* __restrict__ degf, uint8_t* __restrict__ isS, uint8_t* __restrict__ kill) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { isS[i] = degf[i] > 0.0 ? 1 : 0; kill[i] = 0; }
}
// edges (i<j) of As(S,S): the endpoint with the smaller perturbed degree is removed; ties keep
// the smaller index (mis_set.m:49-52)
__global__ void mis_kill_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                const uint8_t* __restrict__ flags, const uint8_t* __restrict__ isS,
                                const double* __restrict__ degf, uint8_t* __restrict__ kill) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n || !isS[row]) return;
    const double di = degf[row];
    bool kill_me = false;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) {
        const int j = idx[e];
        if (flags[e] && j > row && isS[j]) {
            if (di >= degf[j]) kill[j] = 1; else kill_me = true;
        }
    }
    if (kill_me) kill[row] = 1;
}
__global__ void mis_select_kernel(int n, const uint8_t* __restrict__ isS, const uint8_t* __restrict__ kill, uint8_t* __restrict__ isC) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && isS[i] && !kill[i]) isC[i] = 1;
}
// [i,~] = find(As(:,isC)); isF(i) = true (mis_set.m:56-57)
__global__ void mis_markf_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                 const uint8_t* __restrict__ flags, const uint8_t* __restrict__ isC, uint8_t* __restrict__ isF) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    bool hit = false;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) hit |= (flags[e] && isC[idx[e]]);
    if (__any_sync(0xffffffffu, hit) && lane == 0) isF[row] = 1;
}
// isU = ~(isF|isC); deg(~isU) = 0; counts[0] += isC, counts[1] += isU
__global__ void mis_update_kernel(int n, const uint8_t* __restrict__ isC, const uint8_t* __restrict__ isF,
                                  double* __restrict__ degf, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cC = 0, cU = 0;
    if (i < n) {
        const bool u = !(isF[i] || isC[i]);
        if (!u) degf[i] = 0.0;
        cC = isC[i] ? 1 : 0; cU = u ? 1 : 0;
    }
    cC = warp_sum_int(cC); cU = warp_sum_int(cU);
    if ((threadIdx.x & 31) == 0) { if (cC) atomicAdd(counts, cC); if (cU) atomicAdd(counts + 1, cU); }
}
__global__ void mis_leftover_kernel(int n, uint8_t* __restrict__ isC, const uint8_t* __restrict__ isF) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(isF[i] || isC[i])) isC[i] = 1;               // mis_set.m:61-64
}
__global__ void mis_isolated_kernel(int n, const int* __restrict__ rowcnt, uint8_t* __restrict__ isC, uint8_t* __restrict__ isF) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && rowcnt[i] == 0) { isC[i] = 1; isF[i] = 0; }   // mis_set.m:67
}
// ---- all the rounds of mis_set.m:36-67 in ONE launch: a thread-block cluster of 16 CTAs walks init / mark / kill /
// select / mark-F / update round after round with the hardware cluster barrier (release / acquire: global writes of a
// phase are visible to the next) and decides itself when the loop of :42 ends -- 1 launch and no host read instead of
// 5 launches + 1 read per round.  The phases are the kernels above, statement for statement (same flags, same benign
// write races), so isC / isF are the same bit for bit.  counts: 2 ints per round, zeroed by the host.
constexpr int kMisCta = 16;
#ifdef SSN_EMU
constexpr int kMisT = 64;
__device__ __forceinline__ int mis_rank() { return blockIdx.x; }
__device__ __forceinline__ void mis_sync() { emu::cluster_bar->arrive_and_wait(); }
#else
constexpr int kMisT = 1024;
__device__ __forceinline__ int mis_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ void mis_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
#endif
constexpr int kMisMaxRounds = 256;

#ifdef SSN_EMU
void mis_rounds_kernel(
#else
__global__ void __launch_bounds__(kMisT, 1) mis_rounds_kernel(
#endif
        int n, int N0, const int* __restrict__ ptr, const int* __restrict__ idx, const uint8_t* __restrict__ flags,
        const int* __restrict__ deg, const int* __restrict__ rank, const double* __restrict__ rnd, const int* __restrict__ rowcnt,
        double* __restrict__ degf, uint8_t* __restrict__ isS, uint8_t* __restrict__ kill, uint8_t* __restrict__ isC,
        uint8_t* __restrict__ isF, int* __restrict__ counts, int thread_rows) {
    // thread_rows: rows are short (a few entries): one THREAD per row in the two edge passes -- with a warp per row the 512
    // warps of the cluster would walk 32 rows each, one dependent L2 round trip after the other
    const int gt = mis_rank() * kMisT + (int)threadIdx.x, nt = kMisCta * kMisT;
    const int lane = threadIdx.x & 31, gw = gt >> 5, nw = nt >> 5;
    for (int i = gt; i < n; i += nt) {                                   // mis_init_kernel
        const int d = deg[i];
        degf[i] = (d > 0) ? __dadd_rn((double)d, __dmul_rn(0.1, rnd[rank[i]])) : 0.0;
        isC[i] = 0;
        isF[i] = (d == 0) ? 1 : 0;
    }
    mis_sync();
    int sumC = 0, sumU = n, round = 0;
    while ((double)sumC < (double)n / 2.0 && sumU > N0 && round < kMisMaxRounds) {      // mis_set.m:42
        for (int i = gt; i < n; i += nt) { isS[i] = degf[i] > 0.0 ? 1 : 0; kill[i] = 0; }        // mis_mark_kernel
        mis_sync();
        if (thread_rows) {
            for (int row = gt; row < n; row += nt) {                     // mis_kill_kernel, one thread per row
                if (!isS[row]) continue;
                const double di = degf[row];
                bool kill_me = false;
                for (int e = ptr[row]; e < ptr[row + 1]; ++e) {
                    const int j = idx[e];
                    if (flags[e] && j > row && isS[j]) {
                        if (di >= degf[j]) kill[j] = 1; else kill_me = true;
                    }
                }
                if (kill_me) kill[row] = 1;
            }
        } else
        for (int row = gw; row < n; row += nw) {                         // mis_kill_kernel
            if (!isS[row]) continue;
            const double di = degf[row];
            bool kill_me = false;
            for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) {
                const int j = idx[e];
                if (flags[e] && j > row && isS[j]) {
                    if (di >= degf[j]) kill[j] = 1; else kill_me = true;
                }
            }
            if (kill_me) kill[row] = 1;
        }
        mis_sync();
        for (int i = gt; i < n; i += nt) if (isS[i] && !kill[i]) isC[i] = 1;                       // mis_select_kernel
        mis_sync();
        if (thread_rows) {
            for (int row = gt; row < n; row += nt) {                     // mis_markf_kernel, one thread per row
                bool hit = false;
                for (int e = ptr[row]; e < ptr[row + 1]; ++e) hit |= (flags[e] && isC[idx[e]]);
                if (hit) isF[row] = 1;
            }
        } else
        for (int row = gw; row < n; row += nw) {                         // mis_markf_kernel
            bool hit = false;
            for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) hit |= (flags[e] && isC[idx[e]]);
            if (__any_sync(0xffffffffu, hit) && lane == 0) isF[row] = 1;
        }
        mis_sync();
        {                                                                // mis_update_kernel
            int cC = 0, cU = 0;
            for (int i = gt; i < n; i += nt) {
                const bool u = !(isF[i] || isC[i]);
                if (!u) degf[i] = 0.0;
                cC += isC[i] ? 1 : 0; cU += u ? 1 : 0;
            }
            cC = warp_sum_int(cC); cU = warp_sum_int(cU);
            if (lane == 0) { if (cC) atomicAdd(counts + 2 * round, cC); if (cU) atomicAdd(counts + 2 * round + 1, cU); }
        }
        mis_sync();
        sumC = *reinterpret_cast<volatile int*>(counts + 2 * round); sumU = *reinterpret_cast<volatile int*>(counts + 2 * round + 1);
        if (sumU <= N0) {                                                // mis_leftover_kernel   mis_set.m:61-64
            for (int i = gt; i < n; i += nt) if (!(isF[i] || isC[i])) isC[i] = 1;
            sumU = 0;
            mis_sync();
        }
        ++round;
    }
    for (int i = gt; i < n; i += nt) if (rowcnt[i] == 0) { isC[i] = 1; isF[i] = 0; }               // mis_isolated_kernel  :67
}

__global__ void mis_random_kernel(int n0, int n, const double* __restrict__ rnd, uint8_t* __restrict__ isC) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n0) {
        long long pick = (long long)ceil(__dmul_rn(rnd[t], (double)n)) - 1;   // ceil(rand*N), 1-based
        if (pick >= 0 && pick < n) isC[pick] = 1;
    }
}
__global__ void not_kernel(int n, const uint8_t* __restrict__ a, uint8_t* __restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i] ? 0 : 1;
}

}  // namespace

// launches mis_rounds_kernel on one 16-CTA cluster; false when the device cannot co-schedule it (the caller goes round by round)
static bool mis_rounds_cluster(ssn_ctx* c, const CsrView& A, int N0, const uint8_t* flags, const int* deg, const int* rank, const double* rnd,
                               const int* rowcnt, double* degf, uint8_t* isS, uint8_t* kill, uint8_t* isC, uint8_t* isF) {
    Buf<int> counts(c, (size_t)2 * kMisMaxRounds);
    counts.zero();
    const int thread_rows = (A.nrows > 0 && (double)A.nnz / A.nrows <= 16.0) ? 1 : 0;
#ifdef SSN_EMU
    emu_launch_cluster(c, mis_rounds_kernel, kMisCta, kMisT, 0, A.nrows, N0, A.ptr, A.idx, flags, deg, rank, rnd, rowcnt, degf, isS, kill, isC, isF,
                       counts.p, thread_rows);
    return true;
#else
    static int ok16 = -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kMisCta); cfg.blockDim = dim3(kMisT); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kMisCta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (ok16 < 0) {
        ok16 = 0;
        int nclusters = 0;
        if (cudaFuncSetAttribute(mis_rounds_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
            cudaOccupancyMaxActiveClusters(&nclusters, mis_rounds_kernel, &cfg) == cudaSuccess && nclusters >= 1) ok16 = 1;
        (void)cudaGetLastError();
    }
    if (!ok16) return false;
    const int n = A.nrows;
    SSN_CUDA(cudaLaunchKernelEx(&cfg, mis_rounds_kernel, n, N0, A.ptr, A.idx, flags, deg, rank, rnd, rowcnt, degf, isS, kill, isC, isF, counts.p, thread_rows));
    c->launches++;
    return true;
#endif
}

void mis_set(ssn_ctx* c, const CsrView& A, double theta, uint8_t* isC, uint8_t* isF, Buf<uint8_t>& flags) {
    SSN_REQUIRE(A.nrows == A.ncols, SSN_E_NOT_SQUARE, "mis_set: matrix must be square");
    const int n = A.nrows;
    const int N0 = std::min((int)std::sqrt((double)n) + 1, 25);           // mis_set.m:12
    flags.alloc(c, A.nnz);
    Buf<int> deg(c, n), rowcnt(c, n);
    deg.zero();
    strength_flags_full(c, A, theta, flags, deg, rowcnt);                 // mis_set.m:25-29
    const int g = cdiv(n, 256), gw = cdiv((int64_t)n * 32, 256);
    // rank of every connected node among the connected nodes, and their number
    Buf<int> pos(c, n), rank(c, (size_t)n + 1);
    SSN_LAUNCH(c, positive_flag_kernel, g, 256, 0, n, deg.p, pos.p);
    const int64_t nconn = scan_counts_to_ptr(c, pos, rank, n);
    SSN_CUDA(cudaMemsetAsync(isC, 0, n, c->stream));
    SSN_CUDA(cudaMemsetAsync(isF, 0, n, c->stream));
    if ((double)nconn < 0.25 * std::sqrt((double)n)) {                    // mis_set.m:30-34
        Buf<double> rnd(c, N0);
        rng_rand(c, N0, rnd);
        SSN_LAUNCH(c, mis_random_kernel, 1, 32, 0, N0, n, rnd.p, isC);
        SSN_LAUNCH(c, not_kernel, g, 256, 0, n, isC, isF);
        return;
    }
    Buf<double> rnd(c, nconn), degf(c, n);
    rng_rand(c, nconn, rnd);
    Buf<uint8_t> isS(c, n), kill(c, n);
    if (c->mis_cluster && n <= (1 << 16) && A.nnz <= ((int64_t)1 << 21) && mis_rounds_cluster(c, A, N0, flags.p, deg.p, rank.p, rnd.p, rowcnt.p,
                                                                                              degf.p, isS.p, kill.p, isC, isF)) return;
    SSN_LAUNCH(c, mis_init_kernel, g, 256, 0, n, deg.p, rank.p, rnd.p, degf.p, isC, isF);
    Buf<int> counts(c, 2);
    int sumC = 0, sumU = n;
    while ((double)sumC < (double)n / 2.0 && sumU > N0) {                 // mis_set.m:42
        SSN_LAUNCH(c, mis_mark_kernel, g, 256, 0, n, degf.p, isS.p, kill.p);
        SSN_LAUNCH(c, mis_kill_kernel, gw, 256, 0, n, A.ptr, A.idx, flags.p, isS.p, degf.p, kill.p);
        SSN_LAUNCH(c, mis_select_kernel, g, 256, 0, n, isS.p, kill.p, isC);
        SSN_LAUNCH(c, mis_markf_kernel, gw, 256, 0, n, A.ptr, A.idx, flags.p, isC, isF);
        counts.zero();
        SSN_LAUNCH(c, mis_update_kernel, g, 256, 0, n, isC, isF, degf.p, counts.p);
        int h[2]; read_back(c, counts.p, h, 2);
        sumC = h[0]; sumU = h[1];
        if (sumU <= N0) {                                                 // mis_set.m:61-64
            SSN_LAUNCH(c, mis_leftover_kernel, g, 256, 0, n, isC, isF);
            sumU = 0;
        }
    }
    SSN_LAUNCH(c, mis_isolated_kernel, g, 256, 0, n, rowcnt.p, isC, isF);
}

// =================================================================== cf_split (K8b)

namespace {

// state: 0 undecided, 1 C, 2 F.  A node is decided once all lower-numbered neighbours are:
// C iff none of them is C (lexicographically-first MIS == the sequential loop of cf_split.m).
__global__ void cfsplit_round_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                     const double* __restrict__ val, const uint8_t* __restrict__ st_in,
                                     uint8_t* __restrict__ st_out, int* __restrict__ remaining) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t s = st_in[i];
    if (s == 0) {
        bool wait = false, lowerC = false;
        for (int e = ptr[i]; e < ptr[i + 1]; ++e) {
            const int j = idx[e];
            if (j >= i) break;                      // sorted columns: only lower neighbours matter
            if (val[e] == 0.0) continue;
            const uint8_t sj = st_in[j];
            if (sj == 0) wait = true; else if (sj == 1) lowerC = true;
        }
        if (lowerC) s = 2; else if (!wait) s = 1;
        if (s == 0) atomicAdd(remaining, 1);
    }
    st_out[i] = s;
}
__global__ void cfsplit_finish_kernel(int n, const uint8_t* __restrict__ st, uint8_t* __restrict__ indC, uint8_t* __restrict__ indF) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { indC[i] = st[i] == 1; indF[i] = st[i] == 2; }
}

}  // namespace

void cf_split(ssn_ctx* c, const CsrView& S, uint8_t* indC, uint8_t* indF) {
    SSN_REQUIRE(S.nrows == S.ncols, SSN_E_NOT_SQUARE, "cf_split: matrix must be square");
    const int n = S.nrows;
    if (n == 0) return;
    Buf<uint8_t> a(c, n), b(c, n);
    a.zero();
    Buf<int> remaining(c, 1);
    uint8_t* in = a.p; uint8_t* out = b.p;
    const int g = cdiv(n, 256);
    for (int guard = 0; guard <= n; ++guard) {
        int rem = 1;
        for (int rep = 0; rep < 16; ++rep) {        // a lower neighbour marked C makes a node F even
            remaining.zero();                       // while other lower neighbours are undecided
            SSN_LAUNCH(c, cfsplit_round_kernel, g, 256, 0, n, S.ptr, S.idx, S.val, in, out, remaining.p);
            std::swap(in, out);
        }
        rem = read_scalar(c, remaining.p);
        if (rem == 0) break;
    }
    SSN_LAUNCH(c, cfsplit_finish_kernel, g, 256, 0, n, in, indC, indF);
}

// =================================================================== interpolation (K9)

namespace {

__global__ void cf_count_kernel(int n, const uint8_t* __restrict__ isC, const uint8_t* __restrict__ isF,
                                int* __restrict__ cflag, int* __restrict__ fflag, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int ov = 0;
    if (i < n) { cflag[i] = isC[i] ? 1 : 0; fflag[i] = isF[i] ? 1 : 0; ov = (isC[i] && isF[i]) ? 1 : 0; }
    ov = warp_sum_int(ov);
    if ((threadIdx.x & 31) == 0 && ov) atomicAdd(counts, ov);
}
__global__ void bigraph_cf_kernel(int n, int nf, uint8_t* __restrict__ isC, uint8_t* __restrict__ isF) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { isF[i] = i < nf; isC[i] = i >= nf; }
}

// For every F row i (new index fi): W1 entries a_ic/(-d_i) over C columns, and (optionally)
// M entries a_ik/(-d_i) over F columns with k == i or strongly connected (transfer.m:49-51).
// PASS 0 counts the nonzero results, PASS 1 writes them.  One warp per row.
template <int PASS>
__global__ void interp_parts_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                    const double* __restrict__ val, const uint8_t* __restrict__ flags,
                                    const uint8_t* __restrict__ isC, const uint8_t* __restrict__ isF,
                                    const int* __restrict__ cidx, const int* __restrict__ fidx, int want_m,
                                    int* __restrict__ w1cnt, int* __restrict__ mcnt, const int* __restrict__ w1ptr,
                                    int* __restrict__ w1idx, double* __restrict__ w1val, const int* __restrict__ mptr,
                                    int* __restrict__ midx, double* __restrict__ mval, int* __restrict__ notdiag) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n || !isF[row]) return;
    const int fi = fidx[row];
    const int e0 = ptr[row], e1 = ptr[row + 1];
    // diagonal (0 if absent)
    double d = 0.0;
    for (int e = e0 + lane; e < e1; e += 32) if (idx[e] == row) d = val[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, d, o); if (t != 0.0) d = t; }
    const double nd = -d;
    int ow = (PASS == 1) ? w1ptr[fi] : 0, om = (PASS == 1 && want_m) ? mptr[fi] : 0;
    int cw = 0, cm = 0;
    for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        bool kw = false, km = false; double v = 0.0; int j = 0;
        if (e < e1) {
            j = idx[e];
            v = __ddiv_rn(val[e], nd);
            const bool nz = (v != 0.0);
            if (isC[j]) kw = nz;
            else if (isF[j]) {
                if (want_m) km = nz && (j == row || flags[e]);
                else if (j != row && val[e] != 0.0 && notdiag) *notdiag = 1;   // bigraph: Aff must be diagonal
            }
        }
        const unsigned bw = __ballot_sync(0xffffffffu, kw), bm = __ballot_sync(0xffffffffu, km);
        if (PASS == 1) {
            if (kw) { const int pos = ow + __popc(bw & ((1u << lane) - 1u)); w1idx[pos] = cidx[j]; w1val[pos] = v; }
            if (km) { const int pos = om + __popc(bm & ((1u << lane) - 1u)); midx[pos] = fidx[j]; mval[pos] = v; }
            ow += __popc(bw); om += __popc(bm);
        } else { cw += __popc(bw); cm += __popc(bm); }
    }
    if (PASS == 0 && lane == 0) { w1cnt[fi] = cw; if (want_m) mcnt[fi] = cm; }
}

// W = D\W with D = diag(W*ones): row sums accumulated left to right from 0.0 (transfer.m:60-62)
__global__ void row_normalise_kernel(int n, const int* __restrict__ ptr, double* __restrict__ val) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double s = 0.0;
    for (int e = ptr[row]; e < ptr[row + 1]; ++e) s = __dadd_rn(s, __dmul_rn(val[e], 1.0));
    for (int e = ptr[row]; e < ptr[row + 1]; ++e) val[e] = __ddiv_rn(val[e], s);
}

// Pro(p,:) = [W ; I]: F node -> its W row, C node -> unit entry (transfer.m:63)
__global__ void pro_count_kernel(int n, const uint8_t* __restrict__ isF, const int* __restrict__ fidx,
                                 const int* __restrict__ wptr, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    counts[i] = isF[i] ? (wptr[fidx[i] + 1] - wptr[fidx[i]]) : 1;
}
__global__ void pro_fill_kernel(int n, const uint8_t* __restrict__ isF, const int* __restrict__ fidx,
                                const int* __restrict__ cidx, const int* __restrict__ wptr, const int* __restrict__ widx,
                                const double* __restrict__ wval, const int* __restrict__ optr, int* __restrict__ oidx,
                                double* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int o = optr[row];
    if (isF[row]) {
        const int s = wptr[fidx[row]], len = wptr[fidx[row] + 1] - s;
        for (int t = lane; t < len; t += 32) { oidx[o + t] = widx[s + t]; oval[o + t] = wval[s + t]; }
    } else if (lane == 0) { oidx[o] = cidx[row]; oval[o] = 1.0; }
}

__global__ void rowsum_kernel(int n, const int* __restrict__ ptr, const double* __restrict__ val, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    double s = 0.0;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) s += val[e];
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}
__global__ void dinv_kernel(int n, const double* __restrict__ diag, double scale_half, double* __restrict__ dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double inv = __ddiv_rn(1.0, diag[i]);
    dinv[i] = scale_half ? __dmul_rn(0.5, inv) : inv;          // Class_AMG.m:56-57 / :72,:84
}

}  // namespace

AmgOptions resolve_options(const ssn_amg_options* o) {
    AmgOptions r;
    r.retol = 1e-12; r.bigph = 0; r.maxit = 50; r.theta = 0.25; r.smoth = 3; r.cycle = 'v';
    r.isnsp = 0; r.inter = 1; r.fnode = 0; r.guess = nullptr;
    if (!o) {   // nargin == 2 defaults, Class_AMG.m:22-23
        r.retol = 1e-12; r.bigph = 0; r.maxit = 20; r.theta = 0.25; r.smoth = 10; r.cycle = 1; r.isnsp = 1; r.inter = 1;
        return r;
    }
    if (!(o->retol < 0) && o->retol == o->retol) r.retol = o->retol;
    if (o->bigph >= 0) r.bigph = o->bigph;
    if (o->maxit >= 0) r.maxit = o->maxit;
    if (!(o->theta < 0) && o->theta == o->theta) r.theta = o->theta;
    if (o->smoth >= 0) r.smoth = o->smoth;
    if (o->cycle >= 0) r.cycle = o->cycle;
    if (o->isnsp >= 0) r.isnsp = o->isnsp;
    if (o->inter >= 0) r.inter = o->inter;
    r.fnode = o->fnode; r.guess = o->guess_dev;
    return r;
}

void transfer(ssn_ctx* c, const CsrView& A, const AmgOptions& o, int level_J, Csr& Ac, Csr& Pro,
              Buf<uint8_t>* isC_out, Buf<uint8_t>* as_out, Csr* Pt_out) {
    SSN_REQUIRE(A.nrows == A.ncols, SSN_E_NOT_SQUARE, "transfer: matrix must be square");
    const int n = A.nrows;
    const bool bigraph = (level_J == 1 && o.bigph);
    SSN_REQUIRE(o.inter < 2, SSN_E_UNSUPPORTED, "transfer: ideal interpolation (inter = 2) is not supported");
    Buf<uint8_t> isC(c, n), isF(c, n), flags;
    const int g = cdiv(n, 256), gw = cdiv((int64_t)n * 32, 256);
    Phase ph_all(c, bigraph ? "setup.transfer(level1)" : "setup.transfer(mis levels)");
    if (bigraph) {
        SSN_REQUIRE(o.fnode > 0 && o.fnode <= n, SSN_E_BIGPH_FNODE, "amg_options.bigph = 1 requires Nf > 0");
        SSN_LAUNCH(c, bigraph_cf_kernel, g, 256, 0, n, o.fnode, isC.p, isF.p);
        if (as_out) { flags.alloc(c, A.nnz); strength_flags(c, A, o.theta, flags); }
    } else {
        Phase ph(c, "setup.mis_set");
        mis_set(c, A, o.theta, isC, isF, flags);                          // transfer.m:41
    }
    // partition check + index maps
    Buf<int> cflag(c, n), fflag(c, n), cidx(c, (size_t)n + 1), fidx(c, (size_t)n + 1), counts(c, 2);
    counts.zero();
    SSN_LAUNCH(c, cf_count_kernel, g, 256, 0, n, isC.p, isF.p, cflag.p, fflag.p, counts.p);
    scan_counts_async(c, cflag, cidx, n);
    scan_counts_async(c, fflag, fidx, n);
    int h3[3];
    read_ints(c, {cidx.p + n, fidx.p + n, counts.p}, h3);                // one synchronisation for the three
    const int Nc = h3[0], Nf = h3[1], overlap = h3[2];
    SSN_REQUIRE(Nc + Nf == n && overlap == 0, SSN_E_CF_PARTITION,
                "C/F split does not partition the nodes (AMG/transfer.m:46 would index out of range)");
    SSN_REQUIRE(Nf > 0 && Nc > 0, SSN_E_COARSEN_STALL, "coarsening stalled (no F or no C node)");
    // W1 and M
    const int want_m = bigraph ? 0 : 1;
    Buf<int> w1cnt(c, Nf), mcnt(c, Nf), notdiag(c, 1);
    notdiag.zero();
    SSN_LAUNCH(c, interp_parts_kernel<0>, gw, 256, 0, n, A.ptr, A.idx, A.val, flags.p, isC.p, isF.p, cidx.p, fidx.p, want_m,
               w1cnt.p, mcnt.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, notdiag.p);
    // row pointers of W1 and M, then their sizes and the bigraph check with one synchronisation
    Csr W1, M;
    W1.c = c; W1.nrows = Nf; W1.ncols = Nc; W1.ptr.alloc(c, (size_t)Nf + 1);
    scan_counts_async(c, w1cnt, W1.ptr, Nf);
    if (want_m) { M.c = c; M.nrows = Nf; M.ncols = Nf; M.ptr.alloc(c, (size_t)Nf + 1); scan_counts_async(c, mcnt, M.ptr, Nf); }
    read_ints(c, {W1.ptr.p + Nf, want_m ? M.ptr.p + Nf : W1.ptr.p + Nf, notdiag.p}, h3);
    if (bigraph) SSN_REQUIRE(h3[2] == 0, SSN_E_NOT_BIGRAPH, "A(1:Nf,1:Nf) is not diagonal");
    W1.nnz = h3[0]; W1.idx.alloc(c, W1.nnz); W1.val.alloc(c, W1.nnz);
    if (want_m) { M.nnz = h3[1]; M.idx.alloc(c, M.nnz); M.val.alloc(c, M.nnz); }
    SSN_LAUNCH(c, interp_parts_kernel<1>, gw, 256, 0, n, A.ptr, A.idx, A.val, flags.p, isC.p, isF.p, cidx.p, fidx.p, want_m,
               nullptr, nullptr, W1.ptr.p, W1.idx.p, W1.val.p, want_m ? M.ptr.p : nullptr, want_m ? M.idx.p : nullptr,
               want_m ? M.val.p : nullptr, nullptr);
    Csr W;
    if (want_m) {
        Phase ph(c, "setup.interp W2 spgemm+add");
        Csr W2 = spgemm(c, M, W1);                                        // transfer.m:51
        W = sparse_add(c, W1, 0.5, W2);                                   // transfer.m:54-55 (always taken)
    } else {
        W = std::move(W1);
    }
    if (o.isnsp == 1) {                                                   // transfer.m:22-24 / :60-62
        SSN_LAUNCH(c, row_normalise_kernel, cdiv(Nf, 128), 128, 0, Nf, W.ptr.p, W.val.p);
        W = drop_zeros(c, W);
    }
    // Pro
    Buf<int> pcnt(c, n);
    SSN_LAUNCH(c, pro_count_kernel, g, 256, 0, n, isF.p, fidx.p, W.ptr.p, pcnt.p);
    Pro = csr_alloc_known(c, n, Nc, pcnt, W.nnz + Nc);                   // every F row is its W row, every C row one entry
    SSN_LAUNCH(c, pro_fill_kernel, gw, 256, 0, n, isF.p, fidx.p, cidx.p, W.ptr.p, W.idx.p, W.val.p, Pro.ptr.p, Pro.idx.p, Pro.val.p);
    // Ac = (Pro'*A)*Pro                                                  // transfer.m:66
    Csr Pt, T1;
    { Phase ph(c, "setup.transpose Pro"); Pt = transpose(c, Pro); }
    { Phase ph(c, bigraph ? "setup.galerkin L1 Pt*A" : "setup.galerkin Lk Pt*A"); T1 = spgemm(c, Pt, A); }
    { Phase ph(c, bigraph ? "setup.galerkin L1 T1*P" : "setup.galerkin Lk T1*P"); Ac = spgemm(c, T1, Pro); }
    if (isC_out) *isC_out = std::move(isC);
    if (as_out) *as_out = std::move(flags);
    if (Pt_out) *Pt_out = std::move(Pt);                                  // the hierarchy keeps Pro' for the restriction
}

int coarsest_threshold(int64_t N) {
    // 1 + fix(N^(1/3)) with the host libm pow (Class_AMG.m:76); never with CUDA pow
    return 1 + (int)std::pow((double)N, 1.0 / 3.0);
}

void amg_clear(ssn_ctx* c) {
    if (c->hier) {
        cudaStreamSynchronize(c->stream);
        delete c->hier; c->hier = nullptr;
    }
}

static constexpr int kSmallN = 2048;
static constexpr int64_t kSmallNnz = 1 << 17;

// xx_dev: where ones'*A*ones of this level is left on the device (amg_setup reads all levels' sums at once)
static void finish_level(ssn_ctx* c, Level& L, bool bigph_level, int Nf, double* xx_dev) {
    Phase ph(c, "setup.finish_level");
    const int n = L.N = (int)L.A.nrows;
    L.bigph = bigph_level ? 1 : 0; L.Nf = bigph_level ? Nf : 0;
    Buf<double> diag(c, n);
    extract_diag(c, L.A, diag);
    L.dinv.alloc(c, n);
    if (n) SSN_LAUNCH(c, dinv_kernel, cdiv(n, 256), 256, 0, n, diag.p, bigph_level ? 0.0 : 1.0, L.dinv.p);
    L.Axi.alloc(c, n);
    if (n) SSN_LAUNCH(c, rowsum_kernel, cdiv((int64_t)n * 32, 256), 256, 0, n, L.A.ptr.p, L.A.val.p, L.Axi.p);
    dev_sum_async(c, L.Axi, n, xx_dev);
    L.r.alloc(c, n); L.e.alloc(c, n); L.g.alloc(c, n);
}

void amg_setup(ssn_ctx* c, const CsrView& A, const AmgOptions& o, int max_levels) {
    SSN_REQUIRE(A.nrows == A.ncols, SSN_E_NOT_SQUARE, "Class_AMG: matrix must be square");
    if (o.bigph) SSN_REQUIRE(o.fnode > 0, SSN_E_BIGPH_FNODE, "amg_options.bigph = 1 requires Nf > 0");
    amg_clear(c);
    Phase ph_setup(c, "amg_setup total");
    // per-site memory of the sparse products (sparse.cu: spgemm): forgotten every 16 hierarchies, so a site whose rows
    // became short again (or a different problem on the same context) is re-learnt at the price of one failed attempt
    struct SiteScope {
        ssn_ctx* c;
        explicit SiteScope(ssn_ctx* cc) : c(cc) {
            if ((c->spgemm_epoch++ & 15) == 0) std::memset(c->spgemm_big, 0, sizeof(c->spgemm_big));
            c->spgemm_site = 0;
        }
        ~SiteScope() { c->spgemm_site = -1; }
    } site_scope(c);
    std::unique_ptr<Hierarchy> H(new Hierarchy());
    H->smoth = o.smoth;
    H->lv.emplace_back();
    H->lv[0].A = csr_copy(c, A);
    Buf<double> xxd(c, 64);                                             // J < 64 (checked in the loop)
    finish_level(c, H->lv[0], o.bigph != 0, o.fnode, xxd.p);
    const int thr = coarsest_threshold(A.nrows);
    int J = 1;
    // Class_AMG.m:76 coarsens down to the threshold; twogrid_bigph.m:41-47 (max_levels = 2) coarsens exactly once
    while (max_levels > 0 ? (J < max_levels) : (H->lv[J - 1].N > thr)) {
        SSN_REQUIRE(J < 64, SSN_E_COARSEN_STALL, "coarsening stalled");
        // the small levels: all remaining coarsening steps in one kernel (amg_setup_fused.cu)
        if (J >= 2 && fused_small_levels(c, *H, o, max_levels > 0 ? -1 : thr, max_levels > 0 ? max_levels - J : kFusedMaxLevels)) {
            J = (int)H->lv.size();
            continue;
        }
        Level nl;
        Buf<uint8_t> isC;
        transfer(c, H->lv[J - 1].A, o, J, nl.A, nl.P, &isC, nullptr, &nl.Pt);
        SSN_REQUIRE(nl.A.nrows < H->lv[J - 1].N, SSN_E_COARSEN_STALL, "coarsening stalled (no F nodes)");
        H->lv[J - 1].isC = std::move(isC);
        finish_level(c, nl, false, 0, xxd.p + J);
        H->lv.push_back(std::move(nl));
        ++J;
    }
    H->J = J;
    {   // ones'*A_k*ones of every level: one host read for the hierarchy
        double hxx[64];
        read_back(c, xxd.p, hxx, (size_t)J);
        for (int k = 0; k < J; ++k) if (!H->lv[k].xx_known) H->lv[k].xx = hxx[k];
    }
    // tail of small levels handled by the single-block cycle kernel
    int sf = J - 1;
    while (sf > 0 && H->lv[sf - 1].N <= kSmallN && H->lv[sf - 1].A.nnz <= kSmallNnz) --sf;
    if (!(H->lv[sf].N <= kSmallN && H->lv[sf].A.nnz <= kSmallNnz)) sf = J;   // coarsest itself is large (never on this path)
    H->small_from = sf;
    std::vector<LevelDev> hd(J);
    for (int k = 0; k < J; ++k) {
        Level& L = H->lv[k];
        L.pcg.alloc(c, (size_t)5 * L.N);               // ping-pong buffer of the smoother / PCG scratch
        LevelDev d{};
        d.N = L.N; d.Nf = L.Nf; d.bigph = L.bigph;
        d.ap = L.A.ptr.p; d.ai = L.A.idx.p; d.av = L.A.val.p;
        d.pp = L.P.ptr.p; d.pi = L.P.idx.p; d.pv = L.P.val.p;
        d.tp = L.Pt.ptr.p; d.ti = L.Pt.idx.p; d.tv = L.Pt.val.p;
        d.dinv = L.dinv.p; d.Axi = L.Axi.p; d.xx = L.xx;
        d.r = L.r.p; d.e = L.e.p; d.g = L.g.p; d.pcg = L.pcg.p;
        hd[k] = d;
    }
    H->dev.alloc(c, J);
    upload_small(c, H->dev.p, hd.data(), sizeof(LevelDev) * J);
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    H->part.alloc(c, 4096);
    H->scal.alloc(c, 64);
    if (!c->no_cluster) build_cluster_plan(c, *H);
    c->hier = H.release();
}

}  // namespace ssn
