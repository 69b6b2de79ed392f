// amg_setup_fused.cu -- the coarsening of the SMALL levels of the Class_AMG hierarchy (N <= 4096 rows) in ONE kernel.
//
// Below a few thousand rows a coarsening step of AMG/transfer.m:41-66 -- strength filter, the randomised MIS rounds of
// AMG/mis_set.m:25-67, the interpolation W = W1 + 0.5*W2, Pro, Pro', the Galerkin product (Pro'*A)*Pro, the smoother
// data of AMG/Class_AMG.m:84 -- is ~70 kernel launches and ~15 host reads of sizes when it is launched piece by piece
// (amg_setup.cu), i.e. pure launch and round-trip latency: 0.85 ms per level whatever its size.  Here ONE thread-block
// CLUSTER (16 CTAs, hardware cluster barrier) walks all remaining levels, phase by phase: the O(nnz) row loops are
// spread over the cluster's warps, the O(N) scans / counts / the random stream / the bump allocator are REPLICATED in
// every CTA (identical inputs, identical results, no communication), sizes never leave the device, arrays come from an
// arena the host provides, and the host reads the level table back once.  The arithmetic of every step is the arithmetic of the piecewise kernels -- explicit roundings, the
// frozen summation orders of the sparse products (k ascending, multiply then add, no FMA) -- so the hierarchy is the
// same bit for bit (tests/test_gpu_amg.py::test_fused_small_level_setup_equals_piecewise).
#include "amg.cuh"

namespace ssn {

namespace {

constexpr int kFT = 1024;                      // threads per CTA of the fused kernel
constexpr int kFW = kFT / 32;
constexpr int kAccCap = 1024;                  // nonzeros of one product row a warp can hold
constexpr int kBitWords = kFusedMaxN / 32;     // bitmap over the columns of a product row
constexpr int kWarpBytes = kAccCap * 8 + kBitWords * 8;     // accumulator + bitmap + word prefix sums of one warp
constexpr int kDotBytes = kWarpBytes + kAccCap * 12;        // dot flavour: + the staged A row (columns, values)
constexpr int kFusedSmem = 196 * 1024;
constexpr int kSW = kFusedSmem / kWarpBytes;   // warps of a CTA that run the row-accumulate products (each owns an accumulator)
constexpr int kDW = kFusedSmem / kDotBytes;    // warps of a CTA that run the dot-flavour product

// ---- the team: one thread-block cluster on the GPU; ONE block under the host emulation of tests/emu (no clusters there)
#ifdef SSN_EMU
__device__ __forceinline__ int team_rank() { return 0; }
__device__ __forceinline__ int team_size() { return 1; }
__device__ __forceinline__ void team_sync() { __syncthreads(); }
#else
__device__ __forceinline__ int team_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ int team_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return (int)r; }
// release / acquire at cluster scope: global writes of every CTA before the barrier are visible to every CTA after it
__device__ __forceinline__ void team_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
#endif

struct FusedArgs {
    int n0, nnz0;                              // the first small level (built piece by piece)
    const int* ap; const int* ai; const double* av;
    double theta; int isnsp; int thr; int max_levels;   // coarsen while N > thr, at most max_levels new levels
    uint32_t* mt_state;
    unsigned char* arena; unsigned long long arena_bytes;
    FusedLevel* out;                           // max_levels entries
    int* status;                               // [0] status (SSN_OK / error code / kFusedOverflow), [1] new levels, [2..3] draws (lo, hi),
                                               // [4] overflow flag raised by any CTA (zeroed by the host)
    long long* prof;                           // optional: cycles per phase (16 slots), accumulated by thread 0
};

struct Shared {
    unsigned long long front, back;            // bump allocator: permanent arrays grow from the front, temporaries from the back
    int status;
    int itmp[8];
    double dtmp[4];
    int wsum[kFW];
    double wsumd[kFW];
    uint32_t mt[624];
    int mti;
    long long drawn;
};

__device__ __forceinline__ unsigned long long align16(unsigned long long v) { return (v + 15ull) & ~15ull; }

// ---- block-wide helpers (all kFT threads must call)

__device__ int blk_sum_int(Shared& S, int v) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum_int(v);
    __syncthreads();
    if (lane == 0) S.wsum[w] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int i = 0; i < kFW; ++i) t += S.wsum[i];
    return t;
}

__device__ double blk_sum_double(Shared& S, double v) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) S.wsumd[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kFW; ++i) t += S.wsumd[i];
    return t;
}

// out[0] = 0, out[i+1] = in[0] + ... + in[i]; returns the total.  in and out may not alias.
__device__ int blk_scan(Shared& S, const int* in, int* out, int n) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int chunk = (n + kFT - 1) / kFT;
    const int i0 = min(n, tid * chunk), i1 = min(n, i0 + chunk);
    int s = 0;
    for (int i = i0; i < i1; ++i) s += in[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    __syncthreads();
    if (lane == 31) S.wsum[w] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kFW; ++i) { const int v = S.wsum[i]; if (i < w) base += v; total += v; }
    int run = base + incl - s;
    for (int i = i0; i < i1; ++i) { out[i] = run; run += in[i]; }
    if (tid == 0) out[n] = total;
    __syncthreads();
    return total;
}

__device__ void* take(Shared& S, unsigned char* arena, unsigned long long bytes, bool permanent) {
    // called by every thread between barriers with identical arguments: the offsets are replicated, thread 0 commits them
    bytes = align16(bytes ? bytes : 16);
    __syncthreads();
    unsigned long long off;
    if (permanent) { off = S.front; } else { off = S.back - bytes; }
    const bool ok = permanent ? (S.front + bytes <= S.back) : (S.back >= S.front + bytes);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) S.status = kFusedOverflow;
        else if (permanent) S.front += bytes; else S.back -= bytes;
    }
    __syncthreads();
    return ok ? (void*)(arena + off) : nullptr;
}

// genrand_res53 draws of the MATLAB stream into out[0..count) (the body of mt_rand_kernel, amg_setup.cu)
__device__ void blk_rand(Shared& S, uint32_t* st, long long count, double* out) {
    const int tid = threadIdx.x;
    auto twist = [](uint32_t cur, uint32_t nxt, uint32_t far) {
        const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
        return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    };
    uint32_t* mt = S.mt;
    uint32_t* words = reinterpret_cast<uint32_t*>(out);
    const long long total = 2 * count;
    long long produced = 0;
    int mti = S.mti;
    while (produced < total) {
        if (mti >= 624) {
            uint32_t v = 0;
            if (tid < 227) v = twist(mt[tid], mt[tid + 1], mt[tid + 397]);
            __syncthreads();
            if (tid < 227) mt[tid] = v;
            __syncthreads();
            if (tid < 227) v = twist(mt[227 + tid], mt[228 + tid], mt[tid]);
            __syncthreads();
            if (tid < 227) mt[227 + tid] = v;
            __syncthreads();
            if (tid < 170) v = twist(mt[454 + tid], mt[(455 + tid) % 624], mt[227 + tid]);
            __syncthreads();
            if (tid < 170) mt[454 + tid] = v;
            __syncthreads();
            mti = 0;
        }
        const int avail = 624 - mti;
        const long long rem = total - produced;
        const int takew = rem < (long long)avail ? (int)rem : avail;
        for (int t = tid; t < takew; t += kFT) {
            uint32_t y = mt[mti + t];
            y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
            words[produced + t] = y;
        }
        mti += takew; produced += takew;
        __syncthreads();
    }
    if (tid == 0) { S.mti = mti; S.drawn += count; }
    __syncthreads();
    for (long long i = tid; i < count; i += kFT) {
        const uint32_t a = words[2 * i] >> 5, b = words[2 * i + 1] >> 6;
        out[i] = ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
    }
    __syncthreads();
    (void)st;
}

// ---- sparse products, one warp per output row, kSW warps.  A operand: rows given by (start, length) pairs so that
// both CSR matrices and the fixed-stride row segments of an earlier product can be read.

struct RowsView {                              // row r = entries [beg(r), beg(r) + len(r)) of (idx, val)
    const int* ptr; const int* cnt; int stride; const int* idx; const double* val;
    __device__ __forceinline__ int beg(int r) const { return ptr ? ptr[r] : r * stride; }
    __device__ __forceinline__ int len(int r) const { return ptr ? (ptr[r + 1] - ptr[r]) : cnt[r]; }
};

struct WarpAcc { unsigned* bits; int* wbase; double* acc; };

// marks the columns of row `r` of A*B in the warp's bitmap and leaves the word prefix sums; returns the candidate count
__device__ int warp_mark(const WarpAcc& W, const RowsView& A, int r, const int* bp, const int* bi, int nwords) {
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < nwords; t += 32) W.bits[t] = 0u;
    __syncwarp();
    const int a0 = A.beg(r), la = A.len(r);
    for (int t = lane; t < la; t += 32) {
        const int k = A.idx[a0 + t];
        const int b0 = bp[k], b1 = bp[k + 1];
        for (int e = b0; e < b1; ++e) { const int j = bi[e]; atomicOr(&W.bits[j >> 5], 1u << (j & 31)); }
    }
    __syncwarp();
    int run = 0;
    for (int base = 0; base < nwords; base += 32) {
        const int t = base + lane;
        const int c = (t < nwords) ? __popc(W.bits[t]) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (t < nwords) W.wbase[t] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    return run;
}

__device__ __forceinline__ int slot_of(const WarpAcc& W, int j) {
    return W.wbase[j >> 5] + __popc(W.bits[j >> 5] & ((1u << (j & 31)) - 1u));
}

// writes the nonzeros of the accumulated row (columns ascending) to (oidx, oval); returns their number
__device__ int warp_emit(const WarpAcc& W, int nwords, int cand, int* oidx, double* oval) {
    const int lane = threadIdx.x & 31;
    int written = 0;
    // candidates are enumerated word by word: lane t of a batch takes word base+t and walks its set bits
    for (int base = 0; base < nwords; base += 32) {
        const int t = base + lane;
        unsigned bits = (t < nwords) ? W.bits[t] : 0u;
        int s = (t < nwords) ? W.wbase[t] : 0;
        // count this lane's nonzero candidates, then place them after the lower lanes' ones
        int mine = 0;
        { unsigned b = bits; int ss = s; while (b) { b &= b - 1u; mine += (W.acc[ss] != 0.0) ? 1 : 0; ++ss; } }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int pos = written + incl - mine;
        while (bits) {
            const int bit = __ffs(bits) - 1; bits &= bits - 1u;
            const double v = W.acc[s];
            if (v != 0.0) { oidx[pos] = t * 32 + bit; oval[pos] = v; ++pos; }
            ++s;
        }
        written += __shfl_sync(0xffffffffu, incl, 31);
    }
    (void)cand;
    return written;
}

// C = A*B, row-accumulate flavour (A rows short): per output row the A entries are taken in ascending k one at a time,
// the lanes add a_ik*b_kj into the slots of their columns -- the additions of one column happen in ascending k, each
// a multiply followed by an add (no FMA), starting from +0.0, exactly like spgemm() of sparse.cu.
// Output: row r at [r*ostride, ...) of (oidx, oval), ocnt[r] nonzeros (exact zeros dropped).
__device__ void blk_spgemm_acc(int* gflag, const WarpAcc& W, const RowsView& A, int nrows, const int* bp,
                               const int* bi, const double* bv, int ncols, int ostride, int* oidx,
                               double* oval, int* ocnt) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nwords = (ncols + 31) >> 5;
    if (w < kSW) {
        for (int r = team_rank() * kSW + w; r < nrows; r += team_size() * kSW) {
            const int cand = warp_mark(W, A, r, bp, bi, nwords);
            if (cand > kAccCap || cand > ostride) { if (lane == 0) { atomicExch(gflag, 1); ocnt[r] = 0; } continue; }
            for (int t = lane; t < cand; t += 32) W.acc[t] = 0.0;
            __syncwarp();
            const int a0 = A.beg(r), la = A.len(r);
            for (int t = 0; t < la; ++t) {
                const int k = A.idx[a0 + t];
                const double a = A.val[a0 + t];
                const int b0 = bp[k], b1 = bp[k + 1];
                for (int e = b0 + lane; e < b1; e += 32) {
                    const int s = slot_of(W, bi[e]);
                    W.acc[s] = __dadd_rn(W.acc[s], __dmul_rn(a, bv[e]));
                }
                __syncwarp();
            }
            const int nz = warp_emit(W, nwords, cand, oidx + (size_t)r * ostride, oval + (size_t)r * ostride);
            if (lane == 0) ocnt[r] = nz;
            __syncwarp();
        }
    }
    team_sync();
}

// C = A*B, dot flavour (A rows long, B rows short), with Bt = B' given: the pattern of row r is marked as above; every
// candidate column j is owned by one lane, which walks row j of B' (k ascending) and looks k up in row r of A (binary
// search, A's columns ascending): c_rj = sum_k a_rk*b_kj in ascending k, multiply then add from +0.0 -- the same
// additions in the same order as the row-accumulate flavour.
__device__ void blk_spgemm_dot(int* gflag, unsigned char* dsm, const RowsView& A, int nrows, const int* bp,
                               const int* bi, const int* tp, const int* ti,
                               const double* tv, int ncols, int ostride, int* oidx, double* oval, int* ocnt) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nwords = (ncols + 31) >> 5;
    if (w < kDW) {
        unsigned char* base = dsm + (size_t)w * kDotBytes;
        WarpAcc W;
        W.acc = reinterpret_cast<double*>(base);
        W.bits = reinterpret_cast<unsigned*>(base + kAccCap * 8);
        W.wbase = reinterpret_cast<int*>(base + kAccCap * 8 + kBitWords * 4);
        double* sval = reinterpret_cast<double*>(base + kWarpBytes);            // the A row, staged: the binary searches run in
        int* sidx = reinterpret_cast<int*>(base + kWarpBytes + kAccCap * 8);    // shared memory instead of through L2
        for (int r = team_rank() * kDW + w; r < nrows; r += team_size() * kDW) {
            const int a0 = A.beg(r), la = A.len(r);
            const int cand = (la <= kAccCap) ? warp_mark(W, A, r, bp, bi, nwords) : kAccCap + 1;
            if (cand > kAccCap || cand > ostride) { if (lane == 0) { atomicExch(gflag, 1); ocnt[r] = 0; } continue; }
            for (int t = lane; t < la; t += 32) { sidx[t] = A.idx[a0 + t]; sval[t] = A.val[a0 + t]; }
            __syncwarp();
            // every lane computes the values of the candidates it will later emit: word t -> lane t % 32
            for (int wb = 0; wb < nwords; wb += 32) {
                const int t = wb + lane;
                unsigned bits = (t < nwords) ? W.bits[t] : 0u;
                int s = (t < nwords) ? W.wbase[t] : 0;
                while (bits) {
                    const int j = t * 32 + __ffs(bits) - 1; bits &= bits - 1u;
                    double acc = 0.0;
                    for (int e = tp[j]; e < tp[j + 1]; ++e) {
                        const int k = ti[e];
                        int lo = 0, hi = la;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sidx[mid] < k) lo = mid + 1; else hi = mid; }
                        if (lo < la && sidx[lo] == k) acc = __dadd_rn(acc, __dmul_rn(sval[lo], tv[e]));
                    }
                    W.acc[s] = acc;
                    ++s;
                }
            }
            __syncwarp();
            const int nz = warp_emit(W, nwords, cand, oidx + (size_t)r * ostride, oval + (size_t)r * ostride);
            if (lane == 0) ocnt[r] = nz;
            __syncwarp();
        }
    }
    team_sync();
}

#define FPROF(slot) do { if (a.prof && tid == 0 && rk == 0) { const long long now__ = clock64(); a.prof[(slot)] += now__ - t_prof; t_prof = now__; } } while (0)
// the global overflow flag, read by everybody after a team barrier: identical in all CTAs
#define FUSED_CHECK() do { if (S.status == SSN_OK && *(volatile int*)(a.status + 4) != 0) S.status = kFusedOverflow; __syncthreads(); if (S.status != SSN_OK) goto done; } while (0)

__global__ void __launch_bounds__(kFT, 1) fused_levels_kernel(const FusedArgs a) {
    extern __shared__ __align__(16) unsigned char f_dsm[];
    __shared__ Shared S;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int rk = team_rank(), nr = team_size();
    const int gtid = rk * kFT + tid, gthreads = nr * kFT;        // element loops spread over the team
    const int gw = rk * kFW + w, gwarps = nr * kFW;              // warp-per-row loops spread over the team
    long long t_prof = clock64();
    WarpAcc W;
    {
        unsigned char* base = f_dsm + (size_t)(w < kSW ? w : 0) * kWarpBytes;
        W.acc = reinterpret_cast<double*>(base);
        W.bits = reinterpret_cast<unsigned*>(base + kAccCap * 8);
        W.wbase = reinterpret_cast<int*>(base + kAccCap * 8 + kBitWords * 4);
    }
    if (tid == 0) { S.front = 0; S.back = a.arena_bytes & ~15ull; S.status = SSN_OK; S.mti = (int)a.mt_state[624]; S.drawn = 0; }
    for (int i = tid; i < 624; i += kFT) S.mt[i] = a.mt_state[i];
    __syncthreads();

    int n = a.n0;
    const int* ap = a.ap; const int* ai = a.ai; const double* av = a.av;
    int nnz = a.nnz0;
    int built = 0;
    int* gflag = a.status + 4;
    while (n > a.thr && built < a.max_levels) {
        const unsigned long long back0 = S.back;                              // temporaries of this level are released at its end
        __syncthreads();
        // ================= mis_set(A, theta)                                  AMG/mis_set.m:25-67
        const int N0 = min((int)sqrt((double)n) + 1, 25);                     // :12
        double* maxrow = (double*)take(S, a.arena, 8ull * n, false);
        uint8_t* flags = (uint8_t*)take(S, a.arena, (unsigned long long)nnz, false);
        int* deg = (int*)take(S, a.arena, 4ull * n, false);
        int* rowcnt = (int*)take(S, a.arena, 4ull * n, false);
        int* pos = (int*)take(S, a.arena, 4ull * n, false);
        int* rank = (int*)take(S, a.arena, 4ull * (n + 1), false);
        double* degf = (double*)take(S, a.arena, 8ull * n, false);
        uint8_t* isS = (uint8_t*)take(S, a.arena, (unsigned long long)n, false);
        uint8_t* kill = (uint8_t*)take(S, a.arena, (unsigned long long)n, false);
        uint8_t* isF = (uint8_t*)take(S, a.arena, (unsigned long long)n, false);
        uint8_t* isC = (uint8_t*)take(S, a.arena, (unsigned long long)n, true);
        int* cflag = (int*)take(S, a.arena, 4ull * n, false);
        int* fflag = (int*)take(S, a.arena, 4ull * n, false);
        int* cidx = (int*)take(S, a.arena, 4ull * (n + 1), false);
        int* fidx = (int*)take(S, a.arena, 4ull * (n + 1), false);
        if (S.status != SSN_OK) goto done;
        // max_row(i) = max_j (D - A)(i,j), <= 0 -> inf                        strength.m:9-10
        for (int row = gw; row < n; row += gwarps) {
            double mx = 0.0;
            for (int e = ap[row] + lane; e < ap[row + 1]; e += 32) if (ai[e] != row) mx = fmax(mx, -av[e]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) maxrow[row] = (mx <= 0.0) ? INFINITY : mx;
        }
        for (int i = gtid; i < n; i += gthreads) deg[i] = 0;
        team_sync();
        // As = strength(A) >= theta; deg = column counts; rowcnt = row counts   mis_set.m:25-29,67
        for (int row = gw; row < n; row += gwarps) {
            int cnt = 0;
            const double mi = maxrow[row];
            for (int e = ap[row] + lane; e < ap[row + 1]; e += 32) {
                const int j = ai[e]; const double v = av[e];
                uint8_t f = 0;
                if (j != row && v != 0.0) f = (__ddiv_rn(-v, fmin(mi, maxrow[j])) >= a.theta) ? 1 : 0;
                flags[e] = f;
                if (f) { ++cnt; atomicAdd(deg + j, 1); }
            }
            cnt = warp_sum_int(cnt);
            if (lane == 0) rowcnt[row] = cnt;
        }
        team_sync();
        for (int i = gtid; i < n; i += gthreads) { pos[i] = deg[i] > 0 ? 1 : 0; isC[i] = 0; isF[i] = 0; }
        team_sync();
        const int nconn = blk_scan(S, pos, rank, n);                          // replicated: every CTA scans all of pos
        FPROF(0);                                                             // allocation + strength
        if ((double)nconn < 0.25 * sqrt((double)n)) {                         // :30-34
            double* rnd = (double*)take(S, a.arena, 8ull * N0, false);
            if (S.status != SSN_OK) goto done;
            blk_rand(S, a.mt_state, N0, rnd);                                 // replicated: the same draws, the same values written
            team_sync();
            if (gtid < N0) {
                const long long pick = (long long)ceil(__dmul_rn(rnd[gtid], (double)n)) - 1;
                if (pick >= 0 && pick < n) isC[pick] = 1;
            }
            team_sync();
            for (int i = gtid; i < n; i += gthreads) isF[i] = isC[i] ? 0 : 1;
            team_sync();
        } else {
            double* rnd = (double*)take(S, a.arena, 8ull * (nconn > 0 ? nconn : 1), false);
            if (S.status != SSN_OK) goto done;
            blk_rand(S, a.mt_state, nconn, rnd);
            team_sync();
            for (int i = gtid; i < n; i += gthreads) {                        // :35,:40
                const int d = deg[i];
                degf[i] = (d > 0) ? __dadd_rn((double)d, __dmul_rn(0.1, rnd[rank[i]])) : 0.0;
                isF[i] = (d == 0) ? 1 : 0;
            }
            team_sync();
            int sumC = 0, sumU = n;
            while ((double)sumC < (double)n / 2.0 && sumU > N0) {              // :42
                for (int i = gtid; i < n; i += gthreads) { isS[i] = degf[i] > 0.0 ? 1 : 0; kill[i] = 0; }
                team_sync();
                for (int row = gw; row < n; row += gwarps) {                  // :49-52
                    if (!isS[row]) continue;
                    const double di = degf[row];
                    bool kill_me = false;
                    for (int e = ap[row] + lane; e < ap[row + 1]; e += 32) {
                        const int j = ai[e];
                        if (flags[e] && j > row && isS[j]) { if (di >= degf[j]) kill[j] = 1; else kill_me = true; }
                    }
                    if (kill_me) kill[row] = 1;
                }
                team_sync();
                for (int i = gtid; i < n; i += gthreads) if (isS[i] && !kill[i]) isC[i] = 1;
                team_sync();
                for (int row = gw; row < n; row += gwarps) {                  // :56-57
                    bool hit = false;
                    for (int e = ap[row] + lane; e < ap[row + 1]; e += 32) hit |= (flags[e] && isC[ai[e]]);
                    if (__any_sync(0xffffffffu, hit) && lane == 0) isF[row] = 1;
                }
                team_sync();
                for (int i = gtid; i < n; i += gthreads) if (isF[i] || isC[i]) degf[i] = 0.0;
                int cC = 0, cU = 0;                                            // replicated counts over all nodes
                for (int i = tid; i < n; i += kFT) { cC += isC[i] ? 1 : 0; cU += (isF[i] || isC[i]) ? 0 : 1; }
                sumC = blk_sum_int(S, cC); sumU = blk_sum_int(S, cU);
                team_sync();
                if (sumU <= N0) {                                             // :61-64
                    for (int i = gtid; i < n; i += gthreads) if (!(isF[i] || isC[i])) isC[i] = 1;
                    sumU = 0;
                    team_sync();
                }
            }
            for (int i = gtid; i < n; i += gthreads) if (rowcnt[i] == 0) { isC[i] = 1; isF[i] = 0; }   // :67
            team_sync();
        }
        FPROF(1);                                                             // random draws + MIS rounds
        // ================= partition check + index maps                       transfer.m:43-47   (replicated)
        int ov = 0;
        for (int i = tid; i < n; i += kFT) { cflag[i] = isC[i] ? 1 : 0; fflag[i] = isF[i] ? 1 : 0; ov += (isC[i] && isF[i]) ? 1 : 0; }
        const int overlap = blk_sum_int(S, ov);
        const int Nc = blk_scan(S, cflag, cidx, n);
        const int Nf = blk_scan(S, fflag, fidx, n);
        team_sync();
        if (Nc + Nf != n || overlap != 0) { if (tid == 0) S.status = SSN_E_CF_PARTITION; __syncthreads(); goto done; }
        if (Nf <= 0 || Nc <= 0) { if (tid == 0) S.status = SSN_E_COARSEN_STALL; __syncthreads(); goto done; }
        {
        // ================= W1 = -Dff\Afc, M = -Dff\(Aff o (I + As_FF))         transfer.m:49-51
        int* w1cnt = (int*)take(S, a.arena, 4ull * Nf, false);
        int* mcnt = (int*)take(S, a.arena, 4ull * Nf, false);
        int* w1ptr = (int*)take(S, a.arena, 4ull * (Nf + 1), false);
        int* mptr = (int*)take(S, a.arena, 4ull * (Nf + 1), false);
        if (S.status != SSN_OK) goto done;
        for (int row = gw; row < n; row += gwarps) {                          // counts of W1 and M
            if (!isF[row]) continue;
            const int fi = fidx[row];
            const int e0 = ap[row], e1 = ap[row + 1];
            double d = 0.0;
            for (int e = e0 + lane; e < e1; e += 32) if (ai[e] == row) d = av[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, d, o); if (t != 0.0) d = t; }
            const double nd = -d;
            int cw = 0, cm = 0;
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                bool kw = false, km = false;
                if (e < e1) {
                    const int j = ai[e];
                    const bool nz = (__ddiv_rn(av[e], nd) != 0.0);
                    if (isC[j]) kw = nz;
                    else if (isF[j]) km = nz && (j == row || flags[e]);
                }
                cw += __popc(__ballot_sync(0xffffffffu, kw)); cm += __popc(__ballot_sync(0xffffffffu, km));
            }
            if (lane == 0) { w1cnt[fi] = cw; mcnt[fi] = cm; }
        }
        team_sync();
        const int nw1 = blk_scan(S, w1cnt, w1ptr, Nf), nm = blk_scan(S, mcnt, mptr, Nf);
        int* w1idx = (int*)take(S, a.arena, 4ull * nw1, false); double* w1val = (double*)take(S, a.arena, 8ull * nw1, false);
        int* midx = (int*)take(S, a.arena, 4ull * nm, false); double* mval = (double*)take(S, a.arena, 8ull * nm, false);
        if (S.status != SSN_OK) goto done;
        team_sync();
        for (int row = gw; row < n; row += gwarps) {                          // fill
            if (!isF[row]) continue;
            const int fi = fidx[row];
            const int e0 = ap[row], e1 = ap[row + 1];
            double d = 0.0;
            for (int e = e0 + lane; e < e1; e += 32) if (ai[e] == row) d = av[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, d, o); if (t != 0.0) d = t; }
            const double nd = -d;
            int ow = w1ptr[fi], om = mptr[fi];
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                bool kw = false, km = false; double v = 0.0; int j = 0;
                if (e < e1) {
                    j = ai[e];
                    v = __ddiv_rn(av[e], nd);
                    const bool nz = (v != 0.0);
                    if (isC[j]) kw = nz;
                    else if (isF[j]) km = nz && (j == row || flags[e]);
                }
                const unsigned bw = __ballot_sync(0xffffffffu, kw), bm = __ballot_sync(0xffffffffu, km);
                if (kw) { const int p = ow + __popc(bw & ((1u << lane) - 1u)); w1idx[p] = cidx[j]; w1val[p] = v; }
                if (km) { const int p = om + __popc(bm & ((1u << lane) - 1u)); midx[p] = fidx[j]; mval[p] = v; }
                ow += __popc(bw); om += __popc(bm);
            }
        }
        team_sync();
        FPROF(2);                                                             // index maps, W1 and M
        // ================= W2 = M*W1, W = W1 + 0.5*W2                          transfer.m:51,54-55
        const int ws = min(Nc, 256);                                          // row stride of the W2 segments (rows of W2 are short)
        int* w2idx = (int*)take(S, a.arena, 4ull * Nf * ws, false);
        double* w2val = (double*)take(S, a.arena, 8ull * Nf * ws, false);
        int* w2cnt = (int*)take(S, a.arena, 4ull * Nf, false);
        int* wcnt = (int*)take(S, a.arena, 4ull * Nf, false);
        int* wptr = (int*)take(S, a.arena, 4ull * (Nf + 1), false);
        int* zcnt = (int*)take(S, a.arena, 4ull * Nf, false);
        int* zptr = (int*)take(S, a.arena, 4ull * (Nf + 1), false);
        if (S.status != SSN_OK) goto done;
        {
            RowsView Mv{mptr, nullptr, 0, midx, mval};
            blk_spgemm_acc(gflag, W, Mv, Nf, w1ptr, w1idx, w1val, Nc, ws, w2idx, w2val, w2cnt);
        }
        FUSED_CHECK();
        FPROF(3);                                                             // W2 = M*W1
        int* widx = nullptr; double* wval = nullptr;
        for (int fill = 0; fill < 2; ++fill) {
            if (fill == 1) {
                const int nw = blk_scan(S, wcnt, wptr, Nf);
                widx = (int*)take(S, a.arena, 4ull * nw, false); wval = (double*)take(S, a.arena, 8ull * nw, false);
                if (S.status != SSN_OK) goto done;
                team_sync();
            }
            for (int row = gtid; row < Nf; row += gthreads) {                 // sparse_add_kernel: one thread per row
                int ea = w1ptr[row], ea1 = w1ptr[row + 1], eb = row * ws, eb1 = eb + w2cnt[row];
                int o = fill ? wptr[row] : 0, cnt = 0;
                while (ea < ea1 || eb < eb1) {
                    const int ca = (ea < ea1) ? w1idx[ea] : 0x7fffffff;
                    const int cb = (eb < eb1) ? w2idx[eb] : 0x7fffffff;
                    double v; int col;
                    if (ca == cb)     { v = __dadd_rn(w1val[ea], __dmul_rn(0.5, w2val[eb])); col = ca; ++ea; ++eb; }
                    else if (ca < cb) { v = w1val[ea]; col = ca; ++ea; }
                    else              { v = __dmul_rn(0.5, w2val[eb]); col = cb; ++eb; }
                    if (v != 0.0) { if (fill) { widx[o] = col; wval[o] = v; ++o; } ++cnt; }
                }
                if (!fill) wcnt[row] = cnt;
            }
            team_sync();
        }
        // ================= row-normalise (isnsp) and drop the zeros             transfer.m:60-62
        if (a.isnsp == 1) {
            for (int row = gtid; row < Nf; row += gthreads) {
                double s = 0.0;
                for (int e = wptr[row]; e < wptr[row + 1]; ++e) s = __dadd_rn(s, __dmul_rn(wval[e], 1.0));
                int cnt = 0;
                for (int e = wptr[row]; e < wptr[row + 1]; ++e) { wval[e] = __ddiv_rn(wval[e], s); cnt += (wval[e] != 0.0) ? 1 : 0; }
                zcnt[row] = cnt;
            }
        } else {
            for (int row = gtid; row < Nf; row += gthreads) zcnt[row] = wptr[row + 1] - wptr[row];
        }
        team_sync();
        const int nz_w = blk_scan(S, zcnt, zptr, Nf);
        // ================= Pro(p,:) = [W ; I]                                  transfer.m:63
        int* pp = (int*)take(S, a.arena, 4ull * (n + 1), true);
        int* pi = (int*)take(S, a.arena, 4ull * (nz_w + Nc), true);
        double* pv = (double*)take(S, a.arena, 8ull * (nz_w + Nc), true);
        int* pcnt = (int*)take(S, a.arena, 4ull * n, false);
        if (S.status != SSN_OK) goto done;
        team_sync();
        for (int i = gtid; i < n; i += gthreads) pcnt[i] = isF[i] ? zcnt[fidx[i]] : 1;
        team_sync();
        const int nnzP = blk_scan(S, pcnt, pp, n);
        team_sync();
        for (int i = gtid; i < n; i += gthreads) {
            int o = pp[i];
            if (isF[i]) {
                const int fr = fidx[i];
                for (int e = wptr[fr]; e < wptr[fr + 1]; ++e) if (wval[e] != 0.0) { pi[o] = widx[e]; pv[o] = wval[e]; ++o; }
            } else { pi[o] = cidx[i]; pv[o] = 1.0; }
        }
        team_sync();
        FPROF(4);                                                             // W, normalisation, Pro
        // ================= Pt = Pro'  (rows of Pt: entries in ascending original row)
        int* tp = (int*)take(S, a.arena, 4ull * (Nc + 1), true);
        int* ti = (int*)take(S, a.arena, 4ull * nnzP, true);
        double* tv = (double*)take(S, a.arena, 8ull * nnzP, true);
        int* tcnt = (int*)take(S, a.arena, 4ull * Nc, false);
        int* tfill = (int*)take(S, a.arena, 4ull * Nc, false);
        if (S.status != SSN_OK) goto done;
        for (int i = gtid; i < Nc; i += gthreads) { tcnt[i] = 0; tfill[i] = 0; }
        team_sync();
        for (int e = gtid; e < nnzP; e += gthreads) atomicAdd(tcnt + pi[e], 1);
        team_sync();
        blk_scan(S, tcnt, tp, Nc);
        team_sync();
        for (int row = gw; row < n; row += gwarps)
            for (int e = pp[row] + lane; e < pp[row + 1]; e += 32) {
                const int cidx_e = pi[e];
                const int p = tp[cidx_e] + atomicAdd(tfill + cidx_e, 1);
                ti[p] = row; tv[p] = pv[e];
            }
        team_sync();
        for (int r = gtid; r < Nc; r += gthreads) {                            // insertion sort of every (short) row by column
            const int b0 = tp[r], b1 = tp[r + 1];
            for (int x = b0 + 1; x < b1; ++x) {
                const int ki = ti[x]; const double kv = tv[x];
                int y = x - 1;
                while (y >= b0 && ti[y] > ki) { ti[y + 1] = ti[y]; tv[y + 1] = tv[y]; --y; }
                ti[y + 1] = ki; tv[y + 1] = kv;
            }
        }
        team_sync();
        FPROF(5);                                                             // Pro'
        // ================= Ac = (Pro'*A)*Pro                                    transfer.m:66
        const int s1 = min(n, kAccCap), s2 = min(Nc, kAccCap);
        int* t1idx = (int*)take(S, a.arena, 4ull * Nc * s1, false);
        double* t1val = (double*)take(S, a.arena, 8ull * Nc * s1, false);
        int* t1cnt = (int*)take(S, a.arena, 4ull * Nc, false);
        int* acidx = (int*)take(S, a.arena, 4ull * Nc * s2, false);
        double* acval = (double*)take(S, a.arena, 8ull * Nc * s2, false);
        int* accnt = (int*)take(S, a.arena, 4ull * Nc, false);
        if (S.status != SSN_OK) goto done;
        {
            RowsView Ptv{tp, nullptr, 0, ti, tv};
            blk_spgemm_acc(gflag, W, Ptv, Nc, ap, ai, av, n, s1, t1idx, t1val, t1cnt);
        }
        FUSED_CHECK();
        FPROF(6);                                                             // T1 = Pro'*A
        {
            RowsView T1v{nullptr, t1cnt, s1, t1idx, t1val};
            blk_spgemm_dot(gflag, f_dsm, T1v, Nc, pp, pi, tp, ti, tv, Nc, s2, acidx, acval, accnt);
        }
        FUSED_CHECK();
        FPROF(7);                                                             // Ac = T1*Pro
        int* cp = (int*)take(S, a.arena, 4ull * (Nc + 1), true);
        const int nnzC = blk_scan(S, accnt, cp, Nc);
        int* ci = (int*)take(S, a.arena, 4ull * nnzC, true);
        double* cv = (double*)take(S, a.arena, 8ull * nnzC, true);
        double* dinv = (double*)take(S, a.arena, 8ull * Nc, true);
        double* Axi = (double*)take(S, a.arena, 8ull * Nc, true);
        if (S.status != SSN_OK) goto done;
        team_sync();
        for (int row = gw; row < Nc; row += gwarps) {
            const int o = cp[row], len = accnt[row];
            const size_t src = (size_t)row * s2;
            for (int t = lane; t < len; t += 32) { ci[o + t] = acidx[src + t]; cv[o + t] = acval[src + t]; }
        }
        team_sync();
        // ================= smoother data of the new level                      Class_AMG.m:84 ; A*ones, ones'*A*ones
        for (int row = gw; row < Nc; row += gwarps) {
            double s = 0.0, dg = 0.0;
            for (int e = cp[row] + lane; e < cp[row + 1]; e += 32) s += cv[e];
            s = warp_sum(s);
            if (lane == 0) {
                for (int e = cp[row]; e < cp[row + 1]; ++e) if (ci[e] == row) { dg = cv[e]; break; }
                dinv[row] = __dmul_rn(0.5, __ddiv_rn(1.0, dg));
                Axi[row] = s;
            }
        }
        team_sync();
        double xs = 0.0;                                                      // replicated: every CTA sums all of A*ones
        for (int i = tid; i < Nc; i += kFT) xs += Axi[i];
        const double xx = blk_sum_double(S, xs);
        if (tid == 0 && rk == 0) {
            FusedLevel& L = a.out[built];
            L.N = Nc; L.nnzA = nnzC; L.nnzP = nnzP; L.parentN = n; L.xx = xx;
            L.ap = cp; L.ai = ci; L.av = cv; L.pp = pp; L.pi = pi; L.pv = pv; L.tp = tp; L.ti = ti; L.tv = tv;
            L.dinv = dinv; L.Axi = Axi; L.parent_isC = isC;
        }
        FPROF(8);                                                             // compaction + smoother data
        // next level
        n = Nc; nnz = nnzC; ap = cp; ai = ci; av = cv;
        ++built;
        }
        team_sync();
        if (tid == 0) S.back = back0;                                         // release this level's temporaries
        __syncthreads();
    }
done:
    team_sync();
    if (rk == 0) {
        for (int i = tid; i < 624; i += kFT) a.mt_state[i] = S.mt[i];
        if (tid == 0) {
            a.mt_state[624] = (uint32_t)S.mti;
            a.status[0] = S.status; a.status[1] = built;
            a.status[2] = (int)(S.drawn & 0xffffffffll); a.status[3] = (int)(S.drawn >> 32);
        }
    }
}

}  // namespace

// Coarsens H below its last level (N <= kFusedMaxN) in one kernel.  Returns false -- with the hierarchy and the
// random stream untouched -- when the fused path does not apply or ran out of room, so that the caller continues
// piece by piece; throws on the reference's error conditions.
bool fused_small_levels(ssn_ctx* c, Hierarchy& H, const AmgOptions& o, int thr, int max_new_levels) {
    Level& L0 = H.lv.back();
    const int n0 = L0.N;
    const int64_t nnz0 = L0.A.nnz;
    if (!c->fused_setup || n0 > kFusedMaxN || nnz0 > kFusedMaxNnz || n0 <= thr || max_new_levels <= 0) return false;
    if (max_new_levels > kFusedMaxLevels) max_new_levels = kFusedMaxLevels;
    Phase ph(c, "setup.fused_small_levels");
    // arena: the product segments dominate -- rows x min(cols, 1024) x 12 bytes, three live at once -- plus the level itself
    const size_t seg = (size_t)n0 * (size_t)std::min(n0, kAccCap) * 12;
    const size_t arena_bytes = std::min<size_t>((size_t)1 << 30, 3 * seg + 64 * (size_t)nnz0 + 1024 * (size_t)n0 + ((size_t)8 << 20));
    Buf<unsigned char> arena(c, arena_bytes);
    Buf<FusedLevel> out(c, kFusedMaxLevels);
    Buf<int> status(c, 4);
    Buf<uint32_t> mt_save(c, 625);
    SSN_CUDA(cudaMemcpyAsync(mt_save.p, c->mt_state, sizeof(uint32_t) * 625, cudaMemcpyDeviceToDevice, c->stream));
    FusedArgs a{};
    a.n0 = n0; a.nnz0 = (int)nnz0; a.ap = L0.A.ptr.p; a.ai = L0.A.idx.p; a.av = L0.A.val.p;
    a.theta = o.theta; a.isnsp = o.isnsp; a.thr = thr; a.max_levels = max_new_levels;
    a.mt_state = c->mt_state; a.arena = arena.p; a.arena_bytes = arena_bytes; a.out = out.p; a.status = status.p;
    static const bool want_prof = getenv("SSN_FUSED_PROF") != nullptr;
    Buf<long long> prof;
    if (want_prof) { prof.alloc(c, 16); prof.zero(); a.prof = prof.p; }
    const size_t smem = (size_t)kSW * (kBitWords * 4 + kBitWords * 4 + kAccCap * 8);
    static bool attr_set = false;
    if (!attr_set) { SSN_CUDA(cudaFuncSetAttribute(fused_levels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
    SSN_LAUNCH(c, fused_levels_kernel, 1, kFT, smem, a);
    int hs[4];
    read_back(c, status.p, hs, 4);
    if (want_prof) {
        long long hp[16]; read_back(c, prof.p, hp, 16);
        static const char* names[9] = {"alloc+strength", "rand+MIS", "maps+W1+M", "W2=M*W1", "W+norm+Pro", "Pro'", "T1=Pro'*A", "Ac=T1*Pro", "compact+smoother"};
        fprintf(stderr, "fused setup (n0=%d nnz0=%lld, %d levels):", n0, (long long)nnz0, hs[1]);
        for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.0fk", names[i], hp[i] / 1e3);
        fprintf(stderr, " cycles\n");
    }
    if (hs[0] == kFusedOverflow) {                              // no room (a row wider than the accumulator, or the arena): piece by piece
        SSN_CUDA(cudaMemcpyAsync(c->mt_state, mt_save.p, sizeof(uint32_t) * 625, cudaMemcpyDeviceToDevice, c->stream));
        return false;
    }
    SSN_REQUIRE(hs[0] != SSN_E_CF_PARTITION, SSN_E_CF_PARTITION,
                "C/F split does not partition the nodes (AMG/transfer.m:46 would index out of range)");
    SSN_REQUIRE(hs[0] != SSN_E_COARSEN_STALL, SSN_E_COARSEN_STALL, "coarsening stalled (no F or no C node)");
    SSN_REQUIRE(hs[0] == SSN_OK, SSN_E_INVALID, "fused setup kernel failed");
    const int built = hs[1];
    c->rng_drawn += ((int64_t)(uint32_t)hs[2]) | ((int64_t)hs[3] << 32);
    std::vector<FusedLevel> hl((size_t)std::max(built, 1));
    if (built > 0) read_back(c, out.p, hl.data(), (size_t)built);
    for (int k = 0; k < built; ++k) {
        const FusedLevel& F = hl[k];
        Level& parent = H.lv.back();
        parent.isC = Buf<uint8_t>::view(c, F.parent_isC, (size_t)F.parentN);
        Level nl;
        nl.N = F.N; nl.Nf = 0; nl.bigph = 0; nl.xx = F.xx; nl.xx_known = true;
        auto csr_view = [&](Csr& M, int nrows, int ncols, int nnz, int* p, int* i, double* v) {
            M.c = c; M.nrows = nrows; M.ncols = ncols; M.nnz = nnz;
            M.ptr = Buf<int>::view(c, p, (size_t)nrows + 1); M.idx = Buf<int>::view(c, i, (size_t)nnz); M.val = Buf<double>::view(c, v, (size_t)nnz);
        };
        csr_view(nl.A, F.N, F.N, F.nnzA, F.ap, F.ai, F.av);
        csr_view(nl.P, F.parentN, F.N, F.nnzP, F.pp, F.pi, F.pv);
        csr_view(nl.Pt, F.N, F.parentN, F.nnzP, F.tp, F.ti, F.tv);
        nl.dinv = Buf<double>::view(c, F.dinv, (size_t)F.N);
        nl.Axi = Buf<double>::view(c, F.Axi, (size_t)F.N);
        nl.r.alloc(c, F.N); nl.e.alloc(c, F.N); nl.g.alloc(c, F.N);
        H.lv.push_back(std::move(nl));
    }
    H.arenas.push_back(std::move(arena));
    return built > 0;
}

}  // namespace ssn
