// plan_ops.cu -- plan-wide (m x n, column-major, fp64) kernels: Ax, Aty, the fused SsN
// residual, and the active-set compaction that feeds ASAt.  All are HBM-bound streaming
// kernels: 128-bit loads/stores with streaming cache hints, >= 8 independent 16-byte
// requests in flight per lane, deterministic (atomic-free) two-stage reductions.
//
// Tiling of the reduction kernels (Ax, fused residual):
//   a block of 8 warps owns a "group" of 8 strips x 128 rows = 1024 rows and a chunk of
//   <= 512 columns; warp w owns strip w.  A lane holds 4 rows (two double2) of 4 columns per
//   iteration.  Row sums stay in registers for the whole chunk; column partials are reduced
//   over the warp with a 4-value transposing butterfly (6 shuffles per 4 columns) and parked
//   in shared memory until the block combines its 8 strips.  Partials go to
//   rowpart[chunk][m] / colpart[group][n]; a tiny finish kernel sums them in fixed order.
#include "common.cuh"
#include <type_traits>
#include "plan_ops.cuh"
#include "sparse.cuh"

namespace ssn {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kStripRows = 128;
constexpr int kGroupRows = kWarps * kStripRows;   // 1024
constexpr int kMaxChunkCols = 512;
constexpr int kMaxTrials = 8;
constexpr size_t kStageBytes = (size_t)kWarps * 2 * 256 * sizeof(double2);        // per-warp double buffer of 4-column batches: 64 KB per block
constexpr int kStageSmemMax = (int)(kWarps * kMaxChunkCols * sizeof(double) + kStageBytes);

struct PlanArgs {
    const double* x;        // MODE_AX: x ; MODE_PROX: w
    const double* p;
    const double* q;
    const double* lam;      // MODE_PROX: [y1 (n) ; y2 (m)]
    const double* gama;     // optional per-entry upper bound
    double gama_s;
    double inv_tk;
    int64_t m, n;
    int cols_per_chunk, num_chunks, num_groups;
    double* rowpart;        // [num_chunks][m]
    double* colpart;        // [num_groups][n]
    double* scalpart;       // [num_blocks][2]  (norm2, count)
    double* phipart;        // G_PHI: [num_blocks] partial sums of phi .* prox(z) (Class 2: gama = phi array, gama_s = lam(m+n+1))
    double* prox_out;
    double* z_out;
    uint8_t* s_out;
    int want_sums;          // row/col sums wanted
    int stage;              // 1: full strips stream through a per-warp double buffer in shared memory (cp.async)
};

enum { MODE_AX = 0, MODE_PROX = 1 };

// Reduce 4 per-lane values over the warp.  On return the lanes with (lane & 7) == 0 hold the
// full sum of value number  idx = 2*bit4(lane) + bit3(lane).
__device__ __forceinline__ double butterfly4(double v0, double v1, double v2, double v3, int lane) {
    const bool hi = lane & 16;
    double a0 = hi ? v2 : v0, s0 = hi ? v0 : v2;
    double a1 = hi ? v3 : v1, s1 = hi ? v1 : v3;
    a0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool mid = lane & 8;
    double b = mid ? a1 : a0, sb = mid ? a0 : a1;
    b += __shfl_xor_sync(0xffffffffu, sb, 8);
    b += __shfl_xor_sync(0xffffffffu, b, 4);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    return b;
}

// G_PHI (Class2/APD_SsN_Class2.m:124-129): gama = Inf, and the array passed as `gama` is phi, `gama_s` the last dual:
// z = (1/tk)*(w - (Aty(lam) + lam(m+n+1)*phi)); the kernel also sums phi .* prox(z)
enum { G_INF = 0, G_SCALAR = 1, G_VECTOR = 2, G_PHI = 3 };

// element offset of row slot k relative to row slot 0 of the lane
template <bool VEC> __device__ __forceinline__ constexpr int roff(int k) { return VEC ? (64 * (k >> 1) + (k & 1)) : (32 * k); }

// One batch of 4 columns x 4 rows per lane.  `off` is the element offset of (column c, row slot 0).
// FULL: every row and column of the batch is inside the plan (no predicates in the hot path).
// STAGED: the batch was copied into the lane's slots of a staging buffer (stage[(2*cc + h) * 32 + lane]) by cp.async.
// UNITW: p == 1 and q == 1 on the block's rows and columns (every configuration the reference ships): p_i*y1_j + y2_i*q_j is
// y1_j + y2_i and the weighted sums are plain sums bit for bit -- three fp64 operations per entry fewer.
template <int MODE, bool VEC, int GM, bool FULL, bool STAGED = false, bool UNITW = false>
__device__ __forceinline__ void plan_batch(const PlanArgs& a, size_t off, int64_t c, int64_t c1, const bool (&rok)[4],
                                           const double (&pv)[4], const double (&y2v)[4], double (&rs)[4], double& n2,
                                           int& cnt, double (&cs)[4], double& ps, double mu, const double2* stage = nullptr) {
    const size_t m = (size_t)a.m;
    constexpr bool GARR = (GM == G_VECTOR || GM == G_PHI);         // a second plan-sized array rides along
    double v[4][4];
    double g[GARR ? 4 : 1][4];
    // ---- issue all loads of the batch first (8 x 16 B per lane in flight)
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const bool cok = FULL || (c + cc < c1);
        const double* xp = a.x + off + (size_t)cc * m;
        if (STAGED) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double2 t = stage[(2 * cc + h) * 32 + (threadIdx.x & 31)];
                v[cc][2 * h] = t.x; v[cc][2 * h + 1] = t.y;
            }
        } else if (VEC) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double2 t = make_double2(0.0, 0.0);
                if (FULL || (cok && rok[2 * h])) t = __ldcs(reinterpret_cast<const double2*>(xp + roff<VEC>(2 * h)));
                v[cc][2 * h] = t.x; v[cc][2 * h + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[cc][k] = (FULL || (cok && rok[k])) ? __ldcs(xp + roff<VEC>(k)) : 0.0;
        }
        if (MODE == MODE_PROX && GARR) {
            const double* gp = a.gama + off + (size_t)cc * m;
#pragma unroll
            for (int k = 0; k < 4; ++k) g[GARR ? cc : 0][k] = (FULL || (cok && rok[k])) ? __ldcs(gp + roff<VEC>(k)) : 0.0;
        }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const bool cok = FULL || (c + cc < c1);
        const double qj = cok ? __ldg(a.q + c + cc) : 0.0;
        double csum = 0.0;
        if (MODE == MODE_AX) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (UNITW) { csum += v[cc][k]; rs[k] += v[cc][k]; }
                else { csum = fma(v[cc][k], pv[k], csum); rs[k] = fma(v[cc][k], qj, rs[k]); }
            }
        } else {
            const double y1j = cok ? __ldg(a.lam + c + cc) : 0.0;
            double px[4], zz[4]; unsigned char sb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // z = (1/tk) * (w - (p_i*y1_j + y2_i*q_j)), rounded like the reference expression
                // `1/tk*(wk-Aty(lk,p,q))` (mul, mul, add, sub, mul; no FMA contraction).
                double aty = UNITW ? __dadd_rn(y1j, y2v[k]) : __dadd_rn(__dmul_rn(pv[k], y1j), __dmul_rn(y2v[k], qj));
                // Class 2: Htlk = Aty(lk(1:m+n),p,q) + lk(m+n+1)*phi                       APD_SsN_Class2.m:124,138
                if (GM == G_PHI) aty = __dadd_rn(aty, __dmul_rn(mu, g[GARR ? cc : 0][k]));
                const double z = __dmul_rn(a.inv_tk, __dsub_rn(v[cc][k], aty));
                const bool live = FULL || (cok && rok[k]);
                const bool nonneg = live && (z >= 0.0);
                bool act; double pz;
                if (GM == G_INF || GM == G_PHI) { act = nonneg; pz = nonneg ? z : 0.0; } // min(max(0,z),Inf)
                else {
                    const double gm = (GM == G_VECTOR) ? g[GARR ? cc : 0][k] : a.gama_s;
                    const bool below = (z <= gm);
                    act = nonneg && below;
                    pz = nonneg ? (below ? z : gm) : (live ? fmin(0.0, gm) : 0.0);       // min(max(0,z),gama)
                }
                px[k] = pz; zz[k] = z; sb[k] = act ? 1 : 0;
                cnt += act ? 1 : 0;
                // the line-search term of APD_SsN_Class1.m:183-187: ||prox(z)||^2 for gama = Inf (prob < 3);
                // ||z||^2 - ||z - prox(z)||^2 = sum prox(z)*(2z - prox(z)) for finite capacities (prob = 3)
                n2 = (GM == G_INF || GM == G_PHI) ? fma(pz, pz, n2) : fma(pz, __dsub_rn(__dadd_rn(z, z), pz), n2);
                if (GM == G_PHI) ps = fma(g[GARR ? cc : 0][k], pz, ps);                   // phi'*prox(z)
                if (a.want_sums) {
                    if (UNITW) { csum += pz; rs[k] += pz; }
                    else { csum = fma(pz, pv[k], csum); rs[k] = fma(pz, qj, rs[k]); }
                }
            }
            if (a.prox_out || a.z_out || a.s_out) {
                if (cok) {
                    const size_t o = off + (size_t)cc * m;
                    if (VEC) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (FULL || rok[2 * h]) {
                                if (a.prox_out) __stcs(reinterpret_cast<double2*>(a.prox_out + o + roff<VEC>(2 * h)),
                                                       make_double2(px[2 * h], px[2 * h + 1]));
                                if (a.z_out) __stcs(reinterpret_cast<double2*>(a.z_out + o + roff<VEC>(2 * h)),
                                                    make_double2(zz[2 * h], zz[2 * h + 1]));
                                if (a.s_out) *reinterpret_cast<uchar2*>(a.s_out + o + roff<VEC>(2 * h)) =
                                                 make_uchar2(sb[2 * h], sb[2 * h + 1]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (FULL || rok[k]) {
                                if (a.prox_out) a.prox_out[o + roff<VEC>(k)] = px[k];
                                if (a.z_out) a.z_out[o + roff<VEC>(k)] = zz[k];
                                if (a.s_out) a.s_out[o + roff<VEC>(k)] = sb[k];
                            }
                        }
                    }
                }
            }
        }
        cs[cc] = csum;
    }
}

template <int MODE, bool VEC, int GM>
__global__ void __launch_bounds__(kThreads, 2) plan_reduce_kernel(const PlanArgs a) {
    extern __shared__ __align__(16) double colbuf[];   // [kWarps][cols_per_chunk], then (a.stage) [kWarps][2 stages][256] double2
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, group = blockIdx.y;
    const int64_t m = a.m, n = a.n;
    const int cpc = a.cols_per_chunk;
    const int64_t c0 = (int64_t)chunk * cpc;
    const int64_t c1 = (c0 + cpc < n) ? (c0 + cpc) : n;
    const int64_t rbase = ((int64_t)group * kWarps + warp) * kStripRows;
    const int64_t row0 = rbase + (VEC ? 2 * lane : lane);

    bool rok[4];
    double pv[4], y2v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = row0 + roff<VEC>(k);
        rok[k] = r < m;
        pv[k] = rok[k] ? a.p[r] : 0.0;
        y2v[k] = (MODE == MODE_PROX && rok[k]) ? a.lam[n + r] : 0.0;
    }
    double rs[4] = {0.0, 0.0, 0.0, 0.0};
    double n2 = 0.0, ps = 0.0;
    const double mu = (MODE == MODE_PROX && GM == G_PHI) ? a.lam[n + m] : 0.0;       // lk(m+n+1), the multiplier of phi
    int cnt = 0;
    const bool strip_full = (rbase + kStripRows <= m);     // warp-uniform
    size_t off = (size_t)c0 * (size_t)m + (size_t)row0;
    const size_t step = 4 * (size_t)m;

    // unit weights on this block's rows and columns?  (block-uniform; decided from the block's own slices of p and q)
    bool unitw;
    {
        bool u = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) u = u && (!rok[k] || pv[k] == 1.0);
        for (int64_t j = c0 + threadIdx.x; j < c1; j += kThreads) u = u && (a.q[j] == 1.0);
        unitw = __syncthreads_and(u ? 1 : 0) != 0;
    }
    auto run = [&](auto unit_tag) {
        constexpr bool UW = decltype(unit_tag)::value;
        auto park = [&](int64_t c, const double (&cs)[4]) {     // column partials of a batch -> the block's buffer
            if (a.want_sums) {
                const double tot = butterfly4(cs[0], cs[1], cs[2], cs[3], lane);
                const int idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
                if ((lane & 7) == 0 && (c + idx) < c1) colbuf[warp * cpc + (int)(c - c0) + idx] = tot;
            }
        };
        int64_t c = c0;
        constexpr bool kCanStage = VEC && GM != G_VECTOR && GM != G_PHI;
        if (kCanStage && a.stage && strip_full) {
            // Full strips: the complete 4-column batches stream through a per-warp double buffer in shared memory.  The copies of
            // batch b+1 (8 x 16 B per lane, cp.async) are in flight for the whole of the arithmetic of batch b, so a warp keeps 4 KB
            // outstanding all the time instead of only while it waits -- what the register-loading form loses once the per-entry
            // arithmetic is no longer small against the memory latency (fused residual: 0.82 of the streaming rate of a bare
            // read of the plan at the same tiling; tools/micro/ldst256.cu).
            double2* st = reinterpret_cast<double2*>(colbuf + (size_t)kWarps * cpc) + (size_t)warp * 512;
            const int nb = (int)((c1 - c0) >> 2);
            auto issue = [&](int b) {
                const double* xp = a.x + off + (size_t)b * step;
                double2* dst = st + (b & 1) * 256 + lane;
    #pragma unroll
                for (int cc = 0; cc < 4; ++cc)
    #pragma unroll
                    for (int h = 0; h < 2; ++h) cp_async16(dst + (2 * cc + h) * 32, xp + (size_t)cc * m + roff<VEC>(2 * h));
                cp_async_commit();
            };
            if (nb > 0) issue(0);
            for (int b = 0; b < nb; ++b, c += 4) {
                if (b + 1 < nb) { issue(b + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
                double cs[4];
                plan_batch<MODE, VEC, GM, true, true, UW>(a, off + (size_t)b * step, c, c1, rok, pv, y2v, rs, n2, cnt, cs, ps, mu, st + (b & 1) * 256);
                park(c, cs);
            }
            off += (size_t)nb * step;
        }
        for (; c < c1; c += 4, off += step) {
            double cs[4];
            if (strip_full && c + 4 <= c1) plan_batch<MODE, VEC, GM, true, false, UW>(a, off, c, c1, rok, pv, y2v, rs, n2, cnt, cs, ps, mu);
            else                           plan_batch<MODE, VEC, GM, false, false, UW>(a, off, c, c1, rok, pv, y2v, rs, n2, cnt, cs, ps, mu);
            park(c, cs);
        }
    };
    if (unitw) run(std::true_type()); else run(std::false_type());
    if (a.want_sums) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (rok[k]) a.rowpart[(size_t)chunk * (size_t)m + row0 + roff<VEC>(k)] = rs[k];
        __syncthreads();
        const int ncols = (int)(c1 - c0);
        for (int j = threadIdx.x; j < ncols; j += kThreads) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += colbuf[w * cpc + j];
            a.colpart[(size_t)group * (size_t)n + c0 + j] = s;
        }
    }
    if (MODE == MODE_PROX) {
        const double t2 = block_sum(n2, red);
        const double tc = block_sum((double)cnt, red);
        if (threadIdx.x == 0) {
            const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
            a.scalpart[2 * b] = t2; a.scalpart[2 * b + 1] = tc;
        }
        if (GM == G_PHI) {
            const double t3 = block_sum(ps, red);
            if (threadIdx.x == 0) a.phipart[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t3;
        }
    }
}

// y[0:n] = sum_g colpart[g][j] ; y[n:n+m] = sum_c rowpart[c][i] ; scal_out = sum of scalparts
__global__ void plan_finish_kernel(const double* __restrict__ rowpart, const double* __restrict__ colpart,
                                   const double* __restrict__ scalpart, int num_chunks, int num_groups,
                                   int64_t m, int64_t n, int num_blocks, double* __restrict__ y,
                                   double* __restrict__ scal_out) {
    __shared__ double red[32];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (y != nullptr) {
        if (gid < n) {
            double s = 0.0;
            for (int g = 0; g < num_groups; ++g) s += colpart[(size_t)g * n + gid];
            y[gid] = s;
        } else if (gid < n + m) {
            const int64_t i = gid - n;
            double s = 0.0;
            for (int c = 0; c < num_chunks; ++c) s += rowpart[(size_t)c * m + i];
            y[gid] = s;
        }
    }
    if (scal_out != nullptr && blockIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int b = threadIdx.x; b < num_blocks; b += blockDim.x) { s0 += scalpart[2 * b]; s1 += scalpart[2 * b + 1]; }
        s0 = block_sum(s0, red);
        s1 = block_sum(s1, red);
        if (threadIdx.x == 0) { scal_out[0] = s0; scal_out[1] = s1; }
    }
}

// ------------------------------------------------------------------ batched line-search trials
// ||prox((w - Aty(lam_t))/tk)||^2 for NT trial dual vectors lam_t in ONE read of w
// (Class1/APD_SsN_Class1.m:189-207 evaluates them one Aty + prox + norm pass at a time).  The
// per-entry arithmetic of every trial is the single-trial kernel's, operation for operation.
struct TrialArgs {
    const double* w; const double* p; const double* q;
    const double* lamT;     // [NT][n+m]
    const double* gama; double gama_s; double inv_tk;
    int64_t m, n, ldl;      // ldl = n + m
    int cols_per_chunk;
    double* scalpart;       // [num_blocks][NT]
    const int* nonunit;     // device flag: some p_i or q_j differs from 1 (null: assume so)
    int nt_valid;           // trial slots >= nt_valid repeat the last valid trial vector (NT is 1, 2, 4 or 8)
};

// UNITW: p == 1 and q == 1 everywhere (every configuration the reference ships): p_i*y1_j + y2_i*q_j
// is then y1_j + y2_i bit for bit, which takes two of the seven fp64 operations per entry and trial
// off the fp64 pipe -- the pipe that bounds this kernel once NT >= 3.
template <bool VEC, int GM, int NT, bool UNITW>
__device__ __forceinline__ void trials_body(const TrialArgs& a, double* dsm) {
    __shared__ double red[32];
    double (*y2s)[kGroupRows] = reinterpret_cast<double (*)[kGroupRows]>(dsm);       // [NT][1024] row parts
    double* y1s = dsm + NT * kGroupRows;                                              // [NT][cpc]   column parts
    double* qs = y1s + NT * a.cols_per_chunk;                                         // [cpc]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, group = blockIdx.y;
    const int64_t m = a.m, n = a.n;
    const int cpc = a.cols_per_chunk;
    const int64_t c0 = (int64_t)chunk * cpc;
    const int64_t c1 = (c0 + cpc < n) ? (c0 + cpc) : n;
    const int64_t rbase = ((int64_t)group * kWarps + warp) * kStripRows;
    const int lrow0 = warp * kStripRows + (VEC ? 2 * lane : lane);          // row slot 0 inside the group
    const int64_t row0 = (int64_t)group * kGroupRows + lrow0;
    bool rok[4];
    double pv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = row0 + roff<VEC>(k);
        rok[k] = r < m;
        pv[k] = (!UNITW && rok[k]) ? a.p[r] : 0.0;
    }
    // the row parts of the NT trial vectors of this group's rows -> shared memory
    for (int i = threadIdx.x; i < NT * kGroupRows; i += kThreads) {
        const int t = i / kGroupRows, lr = i - t * kGroupRows;
        const int64_t r = (int64_t)group * kGroupRows + lr;
        const int ts = (t < a.nt_valid) ? t : (a.nt_valid - 1);
        y2s[t][lr] = (r < m) ? a.lamT[(size_t)ts * a.ldl + n + r] : 0.0;
    }
    for (int i = threadIdx.x; i < NT * cpc; i += kThreads) {
        const int t = i / cpc, j = i - t * cpc;
        const int ts = (t < a.nt_valid) ? t : (a.nt_valid - 1);
        y1s[i] = (c0 + j < c1) ? a.lamT[(size_t)ts * a.ldl + c0 + j] : 0.0;
    }
    for (int j = threadIdx.x; j < cpc; j += kThreads) qs[j] = (c0 + j < c1) ? a.q[c0 + j] : 0.0;
    __syncthreads();
    double n2[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) n2[t] = 0.0;
    const bool strip_full = (rbase + kStripRows <= m);
    size_t off = (size_t)c0 * (size_t)m + (size_t)row0;
    const size_t step = 2 * (size_t)m;

    auto load_batch = [&](double (&v)[2][4], double (&g)[GM == G_VECTOR ? 2 : 1][4], int64_t c, size_t o) {
        const bool full = strip_full && (c + 2 <= c1);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const bool cok = full || (c + cc < c1);
            const double* xp = a.w + o + (size_t)cc * m;
            if (VEC) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double2 t2 = make_double2(0.0, 0.0);
                    if (full || (cok && rok[2 * h])) t2 = __ldcs(reinterpret_cast<const double2*>(xp + roff<VEC>(2 * h)));
                    v[cc][2 * h] = t2.x; v[cc][2 * h + 1] = t2.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[cc][k] = (full || (cok && rok[k])) ? __ldcs(xp + roff<VEC>(k)) : 0.0;
            }
            if (GM == G_VECTOR) {
                const double* gp = a.gama + o + (size_t)cc * m;
#pragma unroll
                for (int k = 0; k < 4; ++k) g[GM == G_VECTOR ? cc : 0][k] = (full || (cok && rok[k])) ? __ldcs(gp + roff<VEC>(k)) : 0.0;
            }
        }
    };

    // one batch of 2 columns x 4 rows per lane for all NT trials; FULL: no bounds predicates at all
    auto compute_batch = [&](auto full_tag, const double (&v)[2][4], const double (&g)[GM == G_VECTOR ? 2 : 1][4], int64_t c) {
        constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const bool cok = FULL || (c + cc < c1);
            const int jc = (int)(c - c0) + cc;                 // < cpc (cpc is a multiple of 4)
            const double qj = UNITW ? 0.0 : qs[jc];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const double y1j = y1s[t * cpc + jc];
                double y2v[4];
                if (VEC) {
                    const double2 lo = *reinterpret_cast<const double2*>(&y2s[t][lrow0]);
                    const double2 hi = *reinterpret_cast<const double2*>(&y2s[t][lrow0 + 64]);
                    y2v[0] = lo.x; y2v[1] = lo.y; y2v[2] = hi.x; y2v[3] = hi.y;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) y2v[k] = y2s[t][lrow0 + 32 * k];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double aty = UNITW ? __dadd_rn(y1j, y2v[k]) : __dadd_rn(__dmul_rn(pv[k], y1j), __dmul_rn(y2v[k], qj));
                    const double z = __dmul_rn(a.inv_tk, __dsub_rn(v[cc][k], aty));
                    double pz;
                    if (GM == G_INF) {
                        // max(0,z) without a branch or an fp64 compare: clear z when its sign bit is set
                        // (-0.0 and NaN contribute like the compare-based form: 0, resp. NaN-or-0 -> see header)
                        const int hi = __double2hiint(z), lo = __double2loint(z);
                        int keep = ~(hi >> 31);
                        if (!FULL) keep = (cok && rok[k]) ? keep : 0;
                        pz = __hiloint2double(hi & keep, lo & keep);
                    } else {
                        const bool live = FULL || (cok && rok[k]);
                        const bool nonneg = live && (z >= 0.0);
                        const double gm = (GM == G_VECTOR) ? g[GM == G_VECTOR ? cc : 0][k] : a.gama_s;
                        pz = nonneg ? ((z <= gm) ? z : gm) : (live ? fmin(0.0, gm) : 0.0);
                    }
                    n2[t] = (GM == G_INF) ? fma(pz, pz, n2[t]) : fma(pz, __dsub_rn(__dadd_rn(z, z), pz), n2[t]);   // APD_SsN_Class1.m:193-197
                }
            }
        }
    };

    double vcur[2][4], vnxt[2][4];
    double gcur[GM == G_VECTOR ? 2 : 1][4], gnxt[GM == G_VECTOR ? 2 : 1][4];
    if (c0 < c1) load_batch(vcur, gcur, c0, off);
    for (int64_t c = c0; c < c1; c += 2, off += step) {
        if (c + 2 < c1) load_batch(vnxt, gnxt, c + 2, off + step);           // prefetch: overlaps the fp64 work below
        if (strip_full && (c + 2 <= c1)) compute_batch(std::true_type(), vcur, gcur, c);
        else                             compute_batch(std::false_type(), vcur, gcur, c);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { vcur[cc][k] = vnxt[cc][k]; if (GM == G_VECTOR) gcur[GM == G_VECTOR ? cc : 0][k] = gnxt[GM == G_VECTOR ? cc : 0][k]; }
        }
    }
    const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const double tot = block_sum(n2[t], red);
        if (threadIdx.x == 0) a.scalpart[b * NT + t] = tot;
    }
}

template <bool VEC, int GM, int NT>
__global__ void __launch_bounds__(kThreads, 2) plan_trials_kernel(const TrialArgs a) {
    extern __shared__ __align__(16) double trials_dsm[];
    if (a.nonunit != nullptr && *a.nonunit == 0) trials_body<VEC, GM, NT, true>(a, trials_dsm);
    else                                     trials_body<VEC, GM, NT, false>(a, trials_dsm);
}

// ------------------------------------------------------------------ screened line-search trials
// The same quantity for NT <= 32 backtracking steps lam_t = lam + alpha_t*zeta of ONE search
// direction, with per-entry streaming work that does not depend on NT.
// An entry contributes to trial t only if z_t > 0.  w - Aty(lam + alpha*zeta) is linear in alpha up
// to rounding, so if the literal residuals of the FIRST and of the LAST trial of the batch are both
// below -tol (tol = 2^-40 of the operand magnitudes, ~4000x the rounding of the expression) every
// trial of the batch has z_t < 0 there and the entry adds exactly 0 to every norm.  Kernels:
//   screen   streams w once (HBM-bound, ~7 fp64 operations per entry whatever NT is) and writes one
//            bit per entry -- "candidate": survived the screen -- as coalesced 32-bit words
//            [block][word][thread], plus the block's candidate count;
//   compact  turns the bits into ONE list of (column, row) pairs in a fixed order (block, thread,
//            word, bit) after a scan of the block counts;
//   eval     walks the list grid-strided: per candidate one gather of w and the literal expression
//            of every step (lam_t = lam + alpha_t*zeta rounded as trial_vectors_kernel does, then
//            the operations of the single-trial kernel), per-thread accumulators, fixed-order
//            reduction -- deterministic, and balanced over the whole GPU.  The candidates of a
//            late-phase plan sit in the few warps that own its support: evaluating them where
//            they were streamed (in-kernel queue per warp, or one thread per owner) left the rest
//            of the grid idle and ran 2-5x slower than the dense kernel it was meant to beat.
// Only entries that add exactly 0 to every norm are dropped, so the result differs from the dense
// kernel's by the summation order alone (a few ulp).
// max(z,0)^2 is accumulated as (z+|z|)^2 = 4*max(z,0)^2 (exact scaling, undone by the finish kernel).
constexpr int kMaxLinTrials = 32;            // steps per evaluation launch
constexpr int kMaxLinBatch = 256;            // steps per screened batch (one read of w)
struct TrialLinArgs {
    const double* w; const double* p; const double* q;
    const double* lam; const double* zeta;    // [n+m]: [column part ; row part] of lam_old and of the direction
    double inv_tk;
    int64_t m, n;
    int cols_per_chunk, words;                // words = ceil(cols_per_chunk / 8): 32-bit mask words per thread
    unsigned* mask;         // [num_blocks][words][kThreads]
    double* scalpart;       // [evaluation blocks][NT]   (4x the block's share of ||prox||^2)
    double* votepart;       // [num_blocks]              candidate entries of the block
    const int* nonunit;
    double alpha[kMaxLinTrials];              // slots >= the valid count repeat the last valid step
};

template <bool VEC, bool UNITW>
__device__ __forceinline__ void trials_screen_body(const TrialLinArgs& a, double al_f, double al_l, double* dsm) {
    __shared__ double red[32];
    const int cpc = a.cols_per_chunk;
    double* yfc = dsm;                        // [cpc] column parts of the first and of the last trial vector
    double* ylc = yfc + cpc;
    double* qs = ylc + cpc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, group = blockIdx.y;
    const int64_t m = a.m, n = a.n;
    const int64_t c0 = (int64_t)chunk * cpc;
    const int64_t c1 = (c0 + cpc < n) ? (c0 + cpc) : n;
    const int64_t rbase = ((int64_t)group * kWarps + warp) * kStripRows;
    const int64_t row0 = rbase + (VEC ? 2 * lane : lane);
    bool rok[4];
    double pv[4], yfi[4], yli[4];             // weights and first / last trial vector at the lane's four rows
    double mag = 0.0, wmag = 1.0;             // largest |lam|+|alpha_0*zeta| and largest weight of the block
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = row0 + roff<VEC>(k);
        rok[k] = r < m;
        pv[k] = (!UNITW && rok[k]) ? a.p[r] : 0.0;
        const double l = rok[k] ? a.lam[n + r] : 0.0, z = rok[k] ? a.zeta[n + r] : 0.0;
        yfi[k] = __dadd_rn(l, __dmul_rn(al_f, z)); yli[k] = __dadd_rn(l, __dmul_rn(al_l, z));
        mag = fmax(mag, fabs(l) + fabs(al_f * z));
        if (!UNITW) wmag = fmax(wmag, fabs(pv[k]));
    }
    for (int j = threadIdx.x; j < cpc; j += kThreads) {
        const bool ok = c0 + j < c1;
        const double l = ok ? a.lam[c0 + j] : 0.0, z = ok ? a.zeta[c0 + j] : 0.0;
        yfc[j] = __dadd_rn(l, __dmul_rn(al_f, z)); ylc[j] = __dadd_rn(l, __dmul_rn(al_l, z));
        const double qj = ok ? a.q[c0 + j] : 0.0;
        qs[j] = qj;
        mag = fmax(mag, fabs(l) + fabs(al_f * z));
        if (!UNITW) wmag = fmax(wmag, fabs(qj));
    }
    // block-wide maxima (NaN/Inf only widen the margin)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mag = fmax(mag, __shfl_xor_sync(0xffffffffu, mag, o)); wmag = fmax(wmag, __shfl_xor_sync(0xffffffffu, wmag, o));
    }
    if (lane == 0) { red[warp] = mag; red[8 + warp] = wmag; }
    __syncthreads();
    mag = 0.0; wmag = 1.0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) { mag = fmax(mag, red[i]); wmag = fmax(wmag, red[8 + i]); }
    __syncthreads();
    constexpr double kScreen = 9.094947017729282e-13;          // 2^-40
    const double tol0 = kScreen * (2.0 * wmag * mag);

    const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    unsigned* mword = a.mask + b * (size_t)a.words * kThreads + threadIdx.x;
    unsigned bits = 0u;
    int cands = 0;
    const bool strip_full = (rbase + kStripRows <= m);
    size_t off = (size_t)c0 * (size_t)m + (size_t)row0;

    // one batch = 4 columns x 4 rows per lane, all 8 x 16 B loads issued before the first use (as plan_batch);
    // the screen bits of the batch go to bits [16*half + 4*cc + k] of the current mask word (8 columns per word)
    auto batch = [&](auto full_tag, int64_t c, size_t o) {
        constexpr bool FULL = decltype(full_tag)::value;
        double v[4][4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const bool cok = FULL || (c + cc < c1);
            const double* xp = a.w + o + (size_t)cc * m;
            if (VEC) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double2 t2 = make_double2(0.0, 0.0);
                    if (FULL || (cok && rok[2 * h])) t2 = __ldcs(reinterpret_cast<const double2*>(xp + roff<VEC>(2 * h)));
                    v[cc][2 * h] = t2.x; v[cc][2 * h + 1] = t2.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[cc][k] = (FULL || (cok && rok[k])) ? __ldcs(xp + roff<VEC>(k)) : 0.0;
            }
        }
        const int sh = (int)(((c - c0) >> 2) & 1) * 16;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const bool cok = FULL || (c + cc < c1);
            const int jc = (int)(c - c0) + cc;                 // < cpc (cpc is a multiple of 4)
            const double yfj = yfc[jc], ylj = ylc[jc];
            const double qj = UNITW ? 0.0 : qs[jc];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double wv = v[cc][k];
                const bool live = FULL || (cok && rok[k]);
                // literal residual of the first and of the last trial of the batch, plus the margin
                const double af = UNITW ? __dadd_rn(yfj, yfi[k]) : __dadd_rn(__dmul_rn(pv[k], yfj), __dmul_rn(yfi[k], qj));
                const double a_l = UNITW ? __dadd_rn(ylj, yli[k]) : __dadd_rn(__dmul_rn(pv[k], ylj), __dmul_rn(yli[k], qj));
                const double tol = fma(fabs(wv), kScreen, tol0);
                const double uf = (wv - af) + tol, ul = (wv - a_l) + tol;
                // both strictly negative (sign bits set) -> no trial of the batch is active at this entry
                const bool cand = live && ((__double2hiint(uf) & __double2hiint(ul)) >= 0);
                bits |= (cand ? 1u : 0u) << (sh + 4 * cc + k);
            }
        }
    };

    const size_t step = 4 * (size_t)m;
    for (int64_t c = c0; c < c1; c += 4, off += step) {
        if (strip_full && c + 4 <= c1) batch(std::true_type(), c, off);
        else                           batch(std::false_type(), c, off);
        if ((((c - c0) >> 2) & 1) == 1 || c + 4 >= c1) {         // word complete (8 columns) or last batch of the chunk
            mword[(size_t)((c - c0) >> 3) * kThreads] = bits;
            cands += __popc(bits);
            bits = 0u;
        }
    }
    const double tv = block_sum((double)cands, red);
    if (threadIdx.x == 0) a.votepart[b] = tv;
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 2) plan_trials_screen_kernel(const __grid_constant__ TrialLinArgs a) {
    extern __shared__ __align__(16) double trials_lin_dsm[];
    const double al_f = a.alpha[0], al_l = a.alpha[1];                       // first and last step of the batch
    if (a.nonunit != nullptr && *a.nonunit == 0) trials_screen_body<VEC, true>(a, al_f, al_l, trials_lin_dsm);
    else                                     trials_screen_body<VEC, false>(a, al_f, al_l, trials_lin_dsm);
}

// off[b] = number of candidates of the blocks before b (fixed order), off[num_blocks] = total
__global__ void __launch_bounds__(1024) cand_scan_kernel(const double* __restrict__ votepart, int num_blocks,
                                                         unsigned long long* __restrict__ off, double* __restrict__ total_out) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    for (int base = 0; base < num_blocks; base += 1024) {
        const int b = base + threadIdx.x;
        const unsigned long long v = (b < num_blocks) ? (unsigned long long)votepart[b] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long before = carry;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        if (b < num_blocks) off[b] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { off[num_blocks] = carry; total_out[0] = (double)carry; }
}

// list[...] = (column, row) of every candidate, in (block, thread, word, bit) order
template <bool VEC>
__global__ void __launch_bounds__(kThreads) cand_compact_kernel(const TrialLinArgs a, const unsigned long long* __restrict__ off,
                                                                int2* __restrict__ list, unsigned long long cap) {
    __shared__ unsigned wsum[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    if (off[(size_t)gridDim.x * gridDim.y] > cap) return;    // the list does not fit: the host repeats the call with the exact size
    if (off[b + 1] == off[b]) return;         // block-uniform: nothing survived the screen here
    const int cpc = a.cols_per_chunk;
    const int64_t c0 = (int64_t)blockIdx.x * cpc;
    const int64_t c1 = (c0 + cpc < a.n) ? (c0 + cpc) : a.n;
    const int64_t row0 = ((int64_t)blockIdx.y * kWarps + warp) * kStripRows + (VEC ? 2 * lane : lane);
    const unsigned* mword = a.mask + b * (size_t)a.words * kThreads + threadIdx.x;
    const int nwords = (int)((c1 - c0 + 7) >> 3);
    unsigned cnt = 0;
    for (int wi = 0; wi < nwords; ++wi) cnt += __popc(mword[(size_t)wi * kThreads]);
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long pos = off[b] + (incl - cnt);
    for (int w = 0; w < warp; ++w) pos += wsum[w];
    if (cnt == 0) return;
    for (int wi = 0; wi < nwords; ++wi) {
        unsigned bits = mword[(size_t)wi * kThreads];
        while (bits != 0u) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1u;
            const int k = bit & 3, cc = (bit >> 2) & 1;
            const int64_t col = c0 + 8 * wi + 2 * (bit >> 3) + cc;
            const int64_t r = row0 + (VEC ? (64 * (k >> 1) + (k & 1)) : (32 * k));
            list[pos++] = make_int2((int)col, (int)r);
        }
    }
}

// scalpart[block][t] = 4 * sum over the block's share of the list of max(z_t,0)^2  (thread g takes g, g+G, ...)
template <int NT, bool UNITW>
__device__ __forceinline__ void cand_eval_body(const TrialLinArgs& a, const int2* __restrict__ list, const unsigned long long* __restrict__ total_dev,
                                               unsigned long long cap) {
    __shared__ double red[32];
    unsigned long long total = *total_dev;
    if (total > cap) total = 0;               // overflow: the host repeats the call
    if ((unsigned long long)blockIdx.x * blockDim.x >= total) {               // block-uniform: no candidate for this block
        if (threadIdx.x < NT) a.scalpart[(size_t)blockIdx.x * NT + threadIdx.x] = 0.0;
        return;
    }
    const int64_t m = a.m, n = a.n;
    double n2[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) n2[t] = 0.0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int2 cr = list[i];
        const double wv = __ldg(a.w + (size_t)cr.x * (size_t)m + (size_t)cr.y);
        const double lj = __ldg(a.lam + cr.x), zj = __ldg(a.zeta + cr.x);
        const double li = __ldg(a.lam + n + cr.y), zi = __ldg(a.zeta + n + cr.y);
        const double pvv = UNITW ? 0.0 : __ldg(a.p + cr.y), qj = UNITW ? 0.0 : __ldg(a.q + cr.x);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double al = a.alpha[t];
            const double y1 = __dadd_rn(lj, __dmul_rn(al, zj));              // lk_old + delta^ll*zeta
            const double y2 = __dadd_rn(li, __dmul_rn(al, zi));
            const double aty = UNITW ? __dadd_rn(y1, y2) : __dadd_rn(__dmul_rn(pvv, y1), __dmul_rn(y2, qj));
            const double z = __dmul_rn(a.inv_tk, __dsub_rn(wv, aty));
            const double t2 = __dadd_rn(z, fabs(z));                         // 2*max(z,0), exact
            n2[t] = fma(t2, t2, n2[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const double tot = block_sum(n2[t], red);
        if (threadIdx.x == 0) a.scalpart[(size_t)blockIdx.x * NT + t] = tot;
    }
}

template <int NT>
__global__ void __launch_bounds__(kThreads) cand_eval_kernel(const __grid_constant__ TrialLinArgs a, const int2* __restrict__ list,
                                                             const unsigned long long* __restrict__ total_dev, unsigned long long cap) {
    if (a.nonunit != nullptr && *a.nonunit == 0) cand_eval_body<NT, true>(a, list, total_dev, cap);
    else                                     cand_eval_body<NT, false>(a, list, total_dev, cap);
}

// out[t] = scale * sum_b scalpart[b][t] (fixed order) for t = blockIdx.x < nt_out
__global__ void __launch_bounds__(256) trials_lin_finish_kernel(const double* __restrict__ scalpart, int num_blocks, int nt,
                                                                double scale, double* __restrict__ out) {
    __shared__ double red[32];
    const int t = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < num_blocks; b += blockDim.x) s += scalpart[(size_t)b * nt + t];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[t] = scale * s;
}

// nonunit[0] |= 1 if some p_i or q_j differs from 1.0 (nonunit is zeroed by the caller)
__global__ void __launch_bounds__(256) unit_weights_kernel(const double* __restrict__ p, int64_t m, const double* __restrict__ q,
                                                           int64_t n, int* __restrict__ nonunit) {
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m + n; i += (int64_t)gridDim.x * blockDim.x)
        bad |= ((i < m ? p[i] : q[i - m]) != 1.0) ? 1 : 0;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(nonunit, 1);
}

// out[t] = sum_b scalpart[b][t]  (fixed order)
__global__ void __launch_bounds__(256) trials_finish_kernel(const double* __restrict__ scalpart, int num_blocks, int nt,
                                                            double* __restrict__ out) {
    __shared__ double red[32];
    for (int t = 0; t < nt; ++t) {
        double s = 0.0;
        for (int b = threadIdx.x; b < num_blocks; b += blockDim.x) s += scalpart[(size_t)b * nt + t];
        s = block_sum(s, red);
        if (threadIdx.x == 0) out[t] = s;
    }
}

// lamT[t] = lam + alpha[t]*zeta ; f0part: per-block partials of ||lamT[t]||^2 and wlk'lamT[t]   (grid: blocks x trials)
__global__ void __launch_bounds__(256) trial_vectors_kernel(int64_t N, int nt, const double* __restrict__ lam,
                                                            const double* __restrict__ zeta, const double* __restrict__ wlk,
                                                            const double* __restrict__ alpha, double* __restrict__ lamT,
                                                            double* __restrict__ f0part) {
    __shared__ double red[32];
    const int t = blockIdx.y;
    const double al = alpha[t];
    double s2 = 0.0, sw = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = __dadd_rn(lam[i], __dmul_rn(al, zeta[i]));      // lk_old + delta^ll*zeta
        lamT[(size_t)t * N + i] = v;
        s2 = fma(v, v, s2); sw = fma(wlk[i], v, sw);
    }
    s2 = block_sum(s2, red);
    sw = block_sum(sw, red);
    if (threadIdx.x == 0) { f0part[((size_t)blockIdx.x * nt + t) * 2] = s2; f0part[((size_t)blockIdx.x * nt + t) * 2 + 1] = sw; }
}
// out[v] = sum over blocks of f0part[block][v]  (fixed order), one block per value v < 2*nt
__global__ void __launch_bounds__(256) trial_f0_finish_kernel(const double* __restrict__ f0part, int nblocks, int nt,
                                                              double* __restrict__ out /* [nt][2] */) {
    __shared__ double red[32];
    const int t = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += f0part[(size_t)b * 2 * nt + t];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[t] = s;
}

// ------------------------------------------------------------------ fused A-ADMM warm start
// Class1/warmup_class1.m:59-75 as two plan-wide kernels per iteration instead of ~45 separate
// plan-sized passes (4 Ax, 2 Aty, prox and a dozen vector updates in the reference):
//   stage A (:63-67)  reads xk vk wk pik lk2 c, writes dd, accumulates Ax(dd)            (6 r + 1 w)
//   stage B (:70-75)  reads dd xk wk pik lk2, writes xk vk wk pik lk2 in place,
//                     accumulates Ax(vk1) and Ax(xk1)                                    (5 r + 5 w)
// The rank-2 terms Aty(.) are formed on the fly from the (n+m)-vectors; Ax(xk) of the next
// iteration is the Ax(xk1) accumulated here.  Same tiling as plan_reduce_kernel, one column per
// step (6 arrays x 2 x 16 B loads in flight per lane).
struct WarmArgs {
    double* xk; double* vk; double* wk; double* pik; double* lk2; double* dd; const double* c;
    const double* p; const double* q; const double* b;      // b = [r ; l]
    const double* lk1; const double* axk;                  // stage A: h1 = lk1 - (Ax(xk) - b)/bk
    const double* y;                                        // stage B: invAAt(Ax(dd))
    const double* gama; double gama_s;
    double ak, bk, gk, muf, etafk, sgk, etagk, tt;
    int64_t m, n; int cols_per_chunk, num_chunks, num_groups;
    double* rowpart; double* colpart;                       // A: Ax(dd)   B: Ax(vk1)
    double* rowpart2; double* colpart2;                     //             B: Ax(xk1)
    // PHI (Class 2, warmup_class2.m): H = [A I; phi' 0]; b, lk1, axk (= H*uk) and y (= invHHt(H*dd)) have n+m+1 entries
    const double* phi; double* phipart;                     // A: [blocks] phi'*dd_x   B: [blocks][2] phi'*vk1_x, phi'*uk1_x
};

template <bool VEC>
__device__ __forceinline__ void ld4(const double* base, const bool (&rok)[4], bool full, double (&v)[4]) {
    if (VEC) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double2 t = make_double2(0.0, 0.0);
            if (full || rok[2 * h]) t = __ldcs(reinterpret_cast<const double2*>(base + roff<VEC>(2 * h)));
            v[2 * h] = t.x; v[2 * h + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (full || rok[k]) ? __ldcs(base + roff<VEC>(k)) : 0.0;
    }
}
template <bool VEC>
__device__ __forceinline__ void st4(double* base, const bool (&rok)[4], bool full, const double (&v)[4]) {
    if (VEC) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (full || rok[2 * h]) __stcs(reinterpret_cast<double2*>(base + roff<VEC>(2 * h)), make_double2(v[2 * h], v[2 * h + 1]));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (full || rok[k]) __stcs(base + roff<VEC>(k), v[k]);
    }
}

template <int STAGE, bool VEC, int GM, bool PHI = false>
__global__ void __launch_bounds__(kThreads, 2) warm_kernel(const WarmArgs a) {
    extern __shared__ double colbuf[];                 // [2][kWarps][cols_per_chunk]
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, group = blockIdx.y;
    const int64_t m = a.m, n = a.n;
    const int cpc = a.cols_per_chunk;
    const int64_t c0 = (int64_t)chunk * cpc;
    const int64_t c1 = (c0 + cpc < n) ? (c0 + cpc) : n;
    const int64_t rbase = ((int64_t)group * kWarps + warp) * kStripRows;
    const int64_t row0 = rbase + (VEC ? 2 * lane : lane);
    bool rok[4];
    double pv[4], b2[4], ur[4];                         // ur: row part of h1 (stage A) / of y (stage B)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = row0 + roff<VEC>(k);
        rok[k] = r < m;
        pv[k] = rok[k] ? a.p[r] : 0.0;
        b2[k] = rok[k] ? a.b[n + r] : 0.0;
        if (STAGE == 0) ur[k] = rok[k] ? (a.lk1[n + r] - (1.0 / a.bk) * (a.axk[n + r] - b2[k])) : 0.0;
        else            ur[k] = rok[k] ? a.y[n + r] : 0.0;
    }
    double rs[4] = {0.0, 0.0, 0.0, 0.0}, rs2[4] = {0.0, 0.0, 0.0, 0.0};
    const bool full = (rbase + kStripRows <= m);
    size_t off = (size_t)c0 * (size_t)m + (size_t)row0;
    double* colbuf2 = colbuf + (size_t)kWarps * cpc;
    const double ak = a.ak, bk = a.bk, ak2 = a.ak * a.ak;
    // PHI: the last entries of b, of h1 = lk1 - (H*uk - b)/bk (stage A) and of y (stage B) multiply phi
    double bl = 0.0, hl = 0.0, ps = 0.0, ps2 = 0.0;
    if (PHI) {
        bl = a.b[n + m];
        hl = (STAGE == 0) ? (a.lk1[n + m] - (1.0 / a.bk) * (a.axk[n + m] - bl)) : a.y[n + m];
    }
    for (int64_t c = c0; c < c1; ++c, off += (size_t)m) {
        const double qj = __ldg(a.q + c);
        double xk[4], wk[4], pik[4], lk2[4], ph[4] = {0.0, 0.0, 0.0, 0.0};
        ld4<VEC>(a.xk + off, rok, full, xk); ld4<VEC>(a.wk + off, rok, full, wk);
        ld4<VEC>(a.pik + off, rok, full, pik); ld4<VEC>(a.lk2 + off, rok, full, lk2);
        if (PHI) ld4<VEC>(a.phi + off, rok, full, ph);
        double cs = 0.0, cs2 = 0.0;
        if (STAGE == 0) {
            double vk[4], cc[4], dd[4];
            ld4<VEC>(a.vk + off, rok, full, vk); ld4<VEC>(a.c + off, rok, full, cc);
            const double b1j = __ldg(a.b + c);
            const double ucj = __ldg(a.lk1 + c) - (1.0 / bk) * (__ldg(a.axk + c) - b1j);      // h1, column part
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double atb = pv[k] * b1j + b2[k] * qj;                                        // Aty(b)
                double ath = pv[k] * ucj + ur[k] * qj;                                        // Aty(h1)
                if (PHI) { atb += bl * ph[k]; ath += hl * ph[k]; }                            // + b(end)*phi, + hlk(m+n+1)*phi  warmup_class2.m:22,68
                const double wxk = (ak * a.gk * vk[k] + (a.gk + a.muf * ak) * xk[k]) / a.etafk;   // :61
                const double h2 = lk2[k] - (1.0 / bk) * (xk[k] - wk[k]) + (ak / bk) * (-(pik[k] - wk[k]));   // :63
                const double cAw = -atb - wk[k];                                              // :64
                const double cAlk = ath + h2;
                const double d = a.etafk * wxk - ak2 * (cc[k] + cAlk + a.sgk * cAw);          // :65
                dd[k] = (full || rok[k]) ? d : 0.0;
                cs = fma(dd[k], pv[k], cs); rs[k] = fma(dd[k], qj, rs[k]);
                if (PHI) ps = fma(ph[k], dd[k], ps);                                          // phi'*dd, warmup_class2.m:72
            }
            st4<VEC>(a.dd + off, rok, full, dd);
        } else {
            double dd[4], xn[4], vn[4], wn[4], pn[4], ln[4];
            ld4<VEC>(a.dd + off, rok, full, dd);
            double g[4];
            if (GM == G_VECTOR) ld4<VEC>(a.gama + off, rok, full, g);
            const double ycj = __ldg(a.y + c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double aty = pv[k] * ycj + ur[k] * qj;                                        // Aty(invAAt(Ax(dd)))
                if (PHI) aty += hl * ph[k];                                                   // + ff(end)*phi, warmup_class2.m:74
                const double x1 = (dd[k] - aty) / (a.etafk + a.tt);                           // :70
                const double v1 = x1 + (x1 - xk[k]) / ak;                                     // :71
                const double wwk = (ak * pik[k] + wk[k]) / (1.0 + ak);                        // :60
                const double blk2 = lk2[k] + (ak / bk) * (v1 - pik[k]);                       // :72
                const double z = wwk - (ak2 / a.etagk) * (-blk2);                             // :73
                double w1;
                if (GM == G_INF) w1 = fmax(0.0, z);
                else { const double gm = (GM == G_VECTOR) ? g[k] : a.gama_s; w1 = fmin(fmax(0.0, z), gm); }
                const double p1 = w1 + (w1 - wk[k]) / ak;                                     // :74
                const double l1 = lk2[k] + (ak / bk) * (v1 - p1);                             // :75
                const bool live = full || rok[k];
                xn[k] = live ? x1 : 0.0; vn[k] = live ? v1 : 0.0; wn[k] = w1; pn[k] = p1; ln[k] = l1;
                cs = fma(vn[k], pv[k], cs); rs[k] = fma(vn[k], qj, rs[k]);
                cs2 = fma(xn[k], pv[k], cs2); rs2[k] = fma(xn[k], qj, rs2[k]);
                if (PHI) { ps = fma(ph[k], vn[k], ps); ps2 = fma(ph[k], xn[k], ps2); }        // phi'*vk1, phi'*uk1 (x block)
            }
            st4<VEC>(a.xk + off, rok, full, xn); st4<VEC>(a.vk + off, rok, full, vn); st4<VEC>(a.wk + off, rok, full, wn);
            st4<VEC>(a.pik + off, rok, full, pn); st4<VEC>(a.lk2 + off, rok, full, ln);
        }
        cs = warp_sum(cs);
        if (STAGE == 1) cs2 = warp_sum(cs2);
        if (lane == 0) { colbuf[warp * cpc + (int)(c - c0)] = cs; if (STAGE == 1) colbuf2[warp * cpc + (int)(c - c0)] = cs2; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (rok[k]) {
            a.rowpart[(size_t)chunk * (size_t)m + row0 + roff<VEC>(k)] = rs[k];
            if (STAGE == 1) a.rowpart2[(size_t)chunk * (size_t)m + row0 + roff<VEC>(k)] = rs2[k];
        }
    __syncthreads();
    const int ncols = (int)(c1 - c0);
    for (int j = threadIdx.x; j < ncols; j += kThreads) {
        double s = 0.0, s2 = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { s += colbuf[w * cpc + j]; if (STAGE == 1) s2 += colbuf2[w * cpc + j]; }
        a.colpart[(size_t)group * (size_t)n + c0 + j] = s;
        if (STAGE == 1) a.colpart2[(size_t)group * (size_t)n + c0 + j] = s2;
    }
    if (PHI) {
        const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        const double t1 = block_sum(ps, red);
        if (STAGE == 0) { if (threadIdx.x == 0) a.phipart[b] = t1; }
        else {
            const double t2 = block_sum(ps2, red);
            if (threadIdx.x == 0) { a.phipart[2 * b] = t1; a.phipart[2 * b + 1] = t2; }
        }
    }
}

// ------------------------------------------------------------------ fused APD outer-iteration updates
// The plan-wide lines of the outer loop of Class1/APD_SsN_Class1.m around the SsN solve:
//   stage 0 (:125-126)   wk = -c + bk*(xk+ak*vk)/ak^2  and  Ax(xk)                      (3 r + 1 w)
//   stage 1 (:239-254)   xk1 = prox((wk-Aty(lk1))/tk), vk1 = xk1+(xk1-xk)/ak, and in the same pass
//                        Ax(xk1), c'*xk1, ||xk1 - prox(xk1-c-Aty(lk1))||^2               (3 r + 2 w)
// (as torch expressions these are ~40 plan-sized passes per outer iteration).
struct ApdArgs {
    const double* c; const double* xk; const double* vk; const double* wk_in;
    double* wk_out; double* xk1; double* vk1;
    const double* p; const double* q; const double* lam;
    const double* gama; double gama_s;
    double ak, bk, inv_tk;
    int64_t m, n; int cols_per_chunk, num_chunks, num_groups;
    double* rowpart; double* colpart; double* scalpart;
    // PHI (Class 2, APD_SsN_Class2.m:121-122, 231-238): lam has n+m+1 entries, H'lam = Aty(lam(1:n+m)) + lam(n+m+1)*phi
    const double* phi; double* phipart;                     // [blocks] phi'*xk (stage 0) / phi'*xk1 (stage 1)
};

template <int GM>
__device__ __forceinline__ double prox_of(double z, double gm) {
    if (GM == G_INF) return (z >= 0.0) ? z : 0.0;
    return (z >= 0.0) ? ((z <= gm) ? z : gm) : fmin(0.0, gm);
}

template <int STAGE, bool VEC, int GM, bool PHI = false>
__global__ void __launch_bounds__(kThreads, 2) apd_kernel(const ApdArgs a) {
    extern __shared__ double colbuf[];                 // [kWarps][cols_per_chunk]
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, group = blockIdx.y;
    const int64_t m = a.m, n = a.n;
    const int cpc = a.cols_per_chunk;
    const int64_t c0 = (int64_t)chunk * cpc;
    const int64_t c1 = (c0 + cpc < n) ? (c0 + cpc) : n;
    const int64_t rbase = ((int64_t)group * kWarps + warp) * kStripRows;
    const int64_t row0 = rbase + (VEC ? 2 * lane : lane);
    bool rok[4];
    double pv[4], y2v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = row0 + roff<VEC>(k);
        rok[k] = r < m;
        pv[k] = rok[k] ? a.p[r] : 0.0;
        y2v[k] = (STAGE == 1 && rok[k]) ? a.lam[n + r] : 0.0;
    }
    double rs[4] = {0.0, 0.0, 0.0, 0.0};
    double s_cx = 0.0, s_kx = 0.0, ps = 0.0;
    const bool full = (rbase + kStripRows <= m);
    size_t off = (size_t)c0 * (size_t)m + (size_t)row0;
    const double ak = a.ak, ak2 = a.ak * a.ak;
    const double mu = (PHI && STAGE == 1) ? a.lam[n + m] : 0.0;
    for (int64_t c = c0; c < c1; ++c, off += (size_t)m) {
        const double qj = __ldg(a.q + c);
        double cc[4], xk[4], ph[4] = {0.0, 0.0, 0.0, 0.0};
        ld4<VEC>(a.c + off, rok, full, cc); ld4<VEC>(a.xk + off, rok, full, xk);
        if (PHI) ld4<VEC>(a.phi + off, rok, full, ph);
        double cs = 0.0;
        if (STAGE == 0) {
            double vk[4], w[4];
            ld4<VEC>(a.vk + off, rok, full, vk);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // rounded exactly like the reference expression (no FMA contraction): wk decides the active set
                w[k] = __dadd_rn(-cc[k], __ddiv_rn(__dmul_rn(a.bk, __dadd_rn(xk[k], __dmul_rn(ak, vk[k]))), ak2));   // :125
                cs = fma(xk[k], pv[k], cs); rs[k] = fma(xk[k], qj, rs[k]);                      // Ax(xk), :126
                if (PHI) ps = fma(ph[k], xk[k], ps);                                            // phi'*xk, APD_SsN_Class2.m:122
            }
            st4<VEC>(a.wk_out + off, rok, full, w);
        } else {
            double w[4], g[4], x1[4], v1[4];
            ld4<VEC>(a.wk_in + off, rok, full, w);
            if (GM == G_VECTOR) ld4<VEC>(a.gama + off, rok, full, g);
            const double y1j = __ldg(a.lam + c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double gm = (GM == G_VECTOR) ? g[k] : a.gama_s;
                double aty = __dadd_rn(__dmul_rn(pv[k], y1j), __dmul_rn(y2v[k], qj));
                if (PHI) aty = __dadd_rn(aty, __dmul_rn(mu, ph[k]));                            // the fused residual's rounding (plan_batch, G_PHI)
                const double z = __dmul_rn(a.inv_tk, __dsub_rn(w[k], aty));                    // zk, :139 arithmetic
                const bool live = full || rok[k];
                const double xn = live ? prox_of<GM>(z, gm) : 0.0;                              // xk1 = prox(zk), :239
                x1[k] = xn;
                v1[k] = __dadd_rn(xn, __ddiv_rn(__dsub_rn(xn, xk[k]), ak));                     // vk1, :239 (feeds wk: no FMA)
                const double z2 = __dsub_rn(__dsub_rn(xn, cc[k]), aty);                         // xk1 - c - Aty(lk1), :242
                const double dk = live ? (xn - prox_of<GM>(z2, gm)) : 0.0;
                s_kx = fma(dk, dk, s_kx);
                s_cx = fma(cc[k], xn, s_cx);                                                    // c'*xk, :253
                cs = fma(xn, pv[k], cs); rs[k] = fma(xn, qj, rs[k]);                            // Ax(xk1), :241
                if (PHI) ps = fma(ph[k], xn, ps);                                               // phi'*xk1
            }
            st4<VEC>(a.xk1 + off, rok, full, x1); st4<VEC>(a.vk1 + off, rok, full, v1);
        }
        cs = warp_sum(cs);
        if (lane == 0) colbuf[warp * cpc + (int)(c - c0)] = cs;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (rok[k]) a.rowpart[(size_t)chunk * (size_t)m + row0 + roff<VEC>(k)] = rs[k];
    __syncthreads();
    const int ncols = (int)(c1 - c0);
    for (int j = threadIdx.x; j < ncols; j += kThreads) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += colbuf[w * cpc + j];
        a.colpart[(size_t)group * (size_t)n + c0 + j] = s;
    }
    if (STAGE == 1) {
        const double t1 = block_sum(s_cx, red), t2 = block_sum(s_kx, red);
        if (threadIdx.x == 0) {
            const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
            a.scalpart[2 * b] = t1; a.scalpart[2 * b + 1] = t2;
        }
    }
    if (PHI) {
        const double t3 = block_sum(ps, red);
        if (threadIdx.x == 0) a.phipart[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t3;
    }
}

// y = (diag(sg,sg) + A*A') \ x  (invAAt.m:13-20) in one block, no host round trip; np = ||p||^2, nq = ||q||^2
__global__ void __launch_bounds__(1024) invaat_block_kernel(int64_t n, int64_t m, const double* __restrict__ x,
                                                            const double* __restrict__ p, const double* __restrict__ q,
                                                            double sg1, double sg2, double* __restrict__ y) {
    __shared__ double red[32];
    double s_np = 0.0, s_nq = 0.0, s_qv = 0.0, s_pv = 0.0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) { const double pi = p[i]; s_np = fma(pi, pi, s_np); s_pv = fma(pi, x[n + i], s_pv); }
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) { const double qj = q[j]; s_nq = fma(qj, qj, s_nq); s_qv = fma(qj, x[j], s_qv); }
    const double np_ = block_sum(s_np, red), nq = block_sum(s_nq, red), qvn = block_sum(s_qv, red), pvm = block_sum(s_pv, red);
    const double den = sg1 * sg2 + sg1 * nq + sg2 * np_;
    for (int64_t v = threadIdx.x; v < n + m; v += blockDim.x) {
        if (v < n) y[v] = x[v] / (sg1 + np_) + (np_ / (sg1 + np_) * qvn - pvm) * q[v] / den;          // invAAt.m:17
        else       y[v] = x[v] / (sg2 + nq) + (nq / (sg2 + nq) * pvm - qvn) * p[v - n] / den;          // invAAt.m:18
    }
}
// lk1 += (ak/bk) * (Ax(vk1) - b)      (warmup_class1.m:75, first block of lk)
__global__ void warm_lk1_kernel(int64_t N, double coef, const double* __restrict__ av, const double* __restrict__ b, double* __restrict__ lk1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) lk1[i] += coef * (av[i] - b[i]);
}
// finish kernel for two partial sets at once
__global__ void plan_finish2_kernel(const double* __restrict__ rowpart, const double* __restrict__ colpart,
                                    const double* __restrict__ rowpart2, const double* __restrict__ colpart2, int num_chunks,
                                    int num_groups, int64_t m, int64_t n, double* __restrict__ y, double* __restrict__ y2) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) {
        double s = 0.0, s2 = 0.0;
        for (int g = 0; g < num_groups; ++g) { s += colpart[(size_t)g * n + gid]; if (y2) s2 += colpart2[(size_t)g * n + gid]; }
        y[gid] = s; if (y2) y2[gid] = s2;
    } else if (gid < n + m) {
        const int64_t i = gid - n;
        double s = 0.0, s2 = 0.0;
        for (int c = 0; c < num_chunks; ++c) { s += rowpart[(size_t)c * m + i]; if (y2) s2 += rowpart2[(size_t)c * m + i]; }
        y[gid] = s; if (y2) y2[gid] = s2;
    }
}


// ------------------------------------------------------------------ Class 2 (partial OT): slack blocks and (n+m+1)-vectors
// u = [x (m*n) ; y (n) ; z (m)], H = [A I ; phi' 0].  The plan-wide kernels above (PHI variants) handle the x block; the
// slack blocks [y ; z] (N = n + m entries, where H' acts as the identity and wc = 0) and the last entries of the
// (N+1)-vectors are one block of work each, run after the stage kernel and its finish kernel.
struct PotWarmSlack {
    double* us; double* vs; double* ws; double* pis; double* lkBs; double* dds;      // slack blocks of uk vk wk pik lk(N+2:end) dd
    const double* b; double* lkA; double* huk;             // N+1 each: b = [r;l;mu], lk(1:N+1), H*uk
    double* hdd;                                            // N+1: stage A, in: Ax(dd_x) (N entries) -> out: H*dd
    const double* ff;                                       // N+1: stage B, invHHt(H*dd)
    const double* av; const double* au;                     // N each: stage B, Ax(vk1_x) and Ax(uk1_x)
    const double* phipart; int nblocks;
    double ak, bk, gk, muf, etafk, sgk, etagk, tt;
    int N;
};
template <int STAGE>
__global__ void __launch_bounds__(1024) pot_warm_slack_kernel(const PotWarmSlack a) {
    __shared__ double red[32];
    const int N = a.N;
    const double ak = a.ak, bk = a.bk, ak2 = a.ak * a.ak;
    for (int i = threadIdx.x; i < N; i += 1024) {
        const double uk = a.us[i], wk = a.ws[i], pik = a.pis[i], lkB = a.lkBs[i];
        if (STAGE == 0) {
            const double vk = a.vs[i];
            const double hA = a.lkA[i] - (1.0 / bk) * (a.huk[i] - a.b[i]);                    // warmup_class2.m:66, first block
            const double wuk = (ak * a.gk * vk + (a.gk + a.muf * ak) * uk) / a.etafk;         // :64
            const double hB = lkB - (1.0 / bk) * (uk - wk) + (ak / bk) * (-(pik - wk));       // :66, second block
            const double cAw = -a.b[i] - wk;                                                  // :67 (Htb = b(1:n+m) here)
            const double cAlk = hA + hB;                                                      // :68
            const double d = a.etafk * wuk - ak2 * (cAlk + a.sgk * cAw);                      // :69 (wc = 0 here)
            a.dds[i] = d;
            a.hdd[i] += d;                                                                    // :72: Ax(dd_x) + dd_s
        } else {
            const double d = a.dds[i];
            const double u1 = (d - a.ff[i]) / (a.etafk + a.tt);                               // :74
            const double v1 = u1 + (u1 - uk) / ak;                                            // :76
            const double wwk = (ak * pik + wk) / (1.0 + ak);                                  // :63
            const double blk = lkB + (ak / bk) * (v1 - pik);                                  // :78
            const double w1 = fmax(0.0, wwk - (ak2 / a.etagk) * (-blk));                      // :79
            const double p1 = w1 + (w1 - wk) / ak;                                            // :81
            const double l1 = lkB + (ak / bk) * (v1 - p1);                                    // :82
            a.us[i] = u1; a.vs[i] = v1; a.ws[i] = w1; a.pis[i] = p1; a.lkBs[i] = l1;
            a.lkA[i] += (ak / bk) * (a.av[i] + v1 - a.b[i]);                                  // :77, :82
            a.huk[i] = a.au[i] + u1;                                                          // H*uk1: the next iteration's :66
        }
    }
    double s1 = 0.0, s2 = 0.0;
    for (int b = threadIdx.x; b < a.nblocks; b += 1024) {
        if (STAGE == 0) s1 += a.phipart[b];
        else { s1 += a.phipart[2 * b]; s2 += a.phipart[2 * b + 1]; }
    }
    s1 = block_sum(s1, red);
    if (STAGE == 1) s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        if (STAGE == 0) a.hdd[N] = s1;                                                        // phi'*dd_x
        else { a.lkA[N] += (ak / bk) * (s1 - a.b[N]); a.huk[N] = s2; }                        // phi'*vk1_x - mu ; phi'*uk1_x
    }
}

// y = invHHt(v,p,q,sg,phi) = (sg*I + H*H') \ v (Class2/invHHt.m:8-17) in one block, no host round trip;
// l = Ax(phi,p,q) (n+m) and ||phi||^2 are formed once per warm start
__global__ void __launch_bounds__(1024) invhht_block_kernel(int64_t n, int64_t m, const double* __restrict__ v, const double* __restrict__ p,
                                                            const double* __restrict__ q, double sg, const double* __restrict__ l,
                                                            double nphi2, double* __restrict__ y) {
    __shared__ double red[32];
    const int64_t N = n + m;
    const double sg1 = sg + 1.0;                                                              // invAAt(.,p,q,sg+1), :9,:12
    double s_np = 0.0, s_nq = 0.0, s_ql = 0.0, s_pl = 0.0, s_qv = 0.0, s_pv = 0.0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
        const double pi = p[i]; s_np = fma(pi, pi, s_np); s_pl = fma(pi, l[n + i], s_pl); s_pv = fma(pi, v[n + i], s_pv);
    }
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double qj = q[j]; s_nq = fma(qj, qj, s_nq); s_ql = fma(qj, l[j], s_ql); s_qv = fma(qj, v[j], s_qv);
    }
    const double np_ = block_sum(s_np, red), nq = block_sum(s_nq, red);
    const double ql = block_sum(s_ql, red), pl = block_sum(s_pl, red), qv = block_sum(s_qv, red), pvv = block_sum(s_pv, red);
    const double den = sg1 * sg1 + sg1 * nq + sg1 * np_;
    auto inv_aat = [&](double x, int64_t i, double qx, double px) {                           // invAAt.m:17-18
        return (i < n) ? (x / (sg1 + np_) + (np_ / (sg1 + np_) * qx - px) * q[i] / den)
                       : (x / (sg1 + nq) + (nq / (sg1 + nq) * px - qx) * p[i - n] / den);
    };
    double s_lVl = 0.0, s_lVv = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double li = l[i];
        s_lVl = fma(li, inv_aat(li, i, ql, pl), s_lVl);
        s_lVv = fma(li, inv_aat(v[i], i, qv, pvv), s_lVv);
    }
    const double lVl = block_sum(s_lVl, red), lVv = block_sum(s_lVv, red);
    const double s = (sg + nphi2) - lVl;                                                      // :8-9
    const double v2 = v[N];
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double Vl = inv_aat(l[i], i, ql, pl), Vv = inv_aat(v[i], i, qv, pvv);
        y[i] = (s * Vv + lVv * Vl - v2 * Vl) / s;                                             // :14,:17
    }
    if (threadIdx.x == 0) y[N] = (v2 - lVv) / s;                                              // :15,:17
}

// The slack blocks and the (N+1)-vectors of the APD outer iteration of Class 2.
//   stage 0 (APD_SsN_Class2.m:121-122): wk_s = bk*(us+ak*vs)/ak^2, huk = [Ax(xk)+us ; phi'xk], wlk = bk1*(lk-(huk-b)/bk)-b
//   stage 1 (:231-238): us1 = prox((wk_s - lk(1:N))/tk), vs1, huk1 = [Ax(xk1)+us1 ; phi'xk1] and the squared KKT residuals
//           scal[2..4] = ||y1-max(y1-lk(1:n),0)||^2, ||z1-max(z1-lk(n+1:N),0)||^2, ||huk1-b||^2 (scal[0..1] = c'xk1 and the x residual
//           come from the plan-wide kernel)
struct PotApdSlack {
    const double* us; const double* vs; const double* ws_in; double* ws_out; double* us1; double* vs1;
    const double* lk; const double* b; double* axk;         // axk: in Ax(x) (N entries) -> out huk (N+1)
    double* wlk; double* scal;
    const double* phipart; int nblocks;
    double ak, bk, bk1, inv_tk;
    int N, n;
};
template <int STAGE>
__global__ void __launch_bounds__(1024) pot_apd_slack_kernel(const PotApdSlack a) {
    __shared__ double red[32];
    const int N = a.N;
    const double ak = a.ak, ak2 = a.ak * a.ak, inv_bk = 1.0 / a.bk;
    double ky = 0.0, kz = 0.0, kl = 0.0, pp = 0.0;
    for (int b = threadIdx.x; b < a.nblocks; b += 1024) pp += a.phipart[b];
    pp = block_sum(pp, red);                                                                  // phi'*x
    for (int i = threadIdx.x; i <= N; i += 1024) {
        if (STAGE == 0) {
            double h;
            if (i < N) {
                a.ws_out[i] = __ddiv_rn(__dmul_rn(a.bk, __dadd_rn(a.us[i], __dmul_rn(ak, a.vs[i]))), ak2);   // :121 (wc = 0 here)
                h = __dadd_rn(a.axk[i], a.us[i]);
            } else h = pp;
            a.axk[i] = h;
            const double t = __dsub_rn(a.lk[i], __dmul_rn(inv_bk, __dsub_rn(h, a.b[i])));      // :122
            a.wlk[i] = __dsub_rn(__dmul_rn(a.bk1, t), a.b[i]);
        } else {
            double h;
            if (i < N) {
                const double z = __dmul_rn(a.inv_tk, __dsub_rn(a.ws_in[i], a.lk[i]));          // zk on the slack blocks, :124-126
                const double s1 = (z >= 0.0) ? z : 0.0;                                        // uk1 = prox(zk), :231
                a.us1[i] = s1;
                a.vs1[i] = __dadd_rn(s1, __ddiv_rn(__dsub_rn(s1, a.us[i]), ak));
                const double t = s1 - a.lk[i];
                const double d = s1 - ((t >= 0.0) ? t : 0.0);                                  // :234-235
                if (i < a.n) ky = fma(d, d, ky); else kz = fma(d, d, kz);
                h = __dadd_rn(a.axk[i], s1);
            } else h = pp;
            a.axk[i] = h;
            const double e = h - a.b[i];                                                       // :233
            kl = fma(e, e, kl);
        }
    }
    if (STAGE == 1) {
        ky = block_sum(ky, red); kz = block_sum(kz, red); kl = block_sum(kl, red);
        if (threadIdx.x == 0) { a.scal[2] = ky; a.scal[3] = kz; a.scal[4] = kl; }
    }
}

struct Tiling { int groups, chunks, cpc; };

Tiling plan_tiling(ssn_ctx* c, int64_t m, int64_t n, int waves = 2) {
    Tiling t;
    t.groups = cdiv(m, kGroupRows);
    // a slab of less than 512 MB (a rank's share of the 128x128 plan on 8 GPUs: 268 MB, a 50 us launch) streams best as ONE wave
    // of longer blocks (measured, tools/microbench_slab.py: 53 us against 56 with two waves, 61 with four)
    if (waves == 2 && (double)m * (double)n * 8.0 < 512.0 * 1024 * 1024) waves = 1;
    if (c->plan_waves > 0 && waves <= 2) waves = c->plan_waves;  // SSN_PLAN_WAVES (development aid)
    const int64_t target = (int64_t)c->num_sms * 2 * waves;      // `waves` waves at 2 blocks / SM (the reductions: two -- their
                                                                 // partial arrays grow with the number of chunks)
    int64_t chunks = target / t.groups; if (chunks < 1) chunks = 1;
    int64_t cpc = (n + chunks - 1) / chunks;
    cpc = ((cpc + 3) / 4) * 4;
    if (cpc > kMaxChunkCols) cpc = kMaxChunkCols;
    if (cpc < 4) cpc = 4;
    t.cpc = (int)cpc;
    t.chunks = cdiv(n, cpc);
    return t;
}

inline bool vec_ok(const void* ptr, int64_t m) {
    return (m % 2 == 0) && ((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0);
}

// ------------------------------------------------------------------ Aty
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 2) aty_kernel(const double* __restrict__ y, const double* __restrict__ p,
                                                          const double* __restrict__ q, int64_t m, int64_t n,
                                                          int cpc, double* __restrict__ z) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * cpc;
    const int64_t c1 = (c0 + cpc < n) ? (c0 + cpc) : n;
    const int64_t rbase = ((int64_t)blockIdx.y * kWarps + warp) * kStripRows;
    int64_t row[4]; bool rok[4]; double pv[4], y2v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        row[k] = VEC ? (rbase + 64 * (k >> 1) + 2 * lane + (k & 1)) : (rbase + 32 * k + lane);
        rok[k] = row[k] < m;
        pv[k] = rok[k] ? p[row[k]] : 0.0;
        y2v[k] = rok[k] ? y[n + row[k]] : 0.0;
    }
    if (!rok[0]) return;                                   // whole lane beyond the slab
    for (int64_t c = c0; c < c1; ++c) {
        const double y1j = __ldg(y + c), qj = __ldg(q + c);
        double o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)                       // Aty.m:12-13: p*y1' , y2*q' , add
            o[k] = __dadd_rn(__dmul_rn(pv[k], y1j), __dmul_rn(y2v[k], qj));
        double* base = z + (size_t)c * (size_t)m;
        if (VEC) {
            __stcs(reinterpret_cast<double2*>(base + row[0]), make_double2(o[0], o[1]));
            if (rok[2]) __stcs(reinterpret_cast<double2*>(base + row[2]), make_double2(o[2], o[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (rok[k]) __stcs(base + row[k], o[k]);
        }
    }
}

// ------------------------------------------------------------------ active-set compaction
// One warp per column.  Pass 1 counts the active rows of each column, pass 2 (after a scan)
// writes them in ascending row order with ballot-free packed-mask prefix sums.
template <bool VEC16>
__device__ __forceinline__ unsigned load_mask16(const uint8_t* col, int64_t i0, int64_t m) {
    unsigned mask = 0;
    if (VEC16) {
        if (i0 + 16 <= m) {
            const uint4 t = __ldcs(reinterpret_cast<const uint4*>(col + i0));
            const unsigned w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if ((w[k] >> (8 * b)) & 0xffu) mask |= 1u << (4 * k + b);
            }
            return mask;
        }
    }
#pragma unroll 4
    for (int b = 0; b < 16; ++b)
        if (i0 + b < m && col[i0 + b]) mask |= 1u << b;
    return mask;
}

template <bool VEC16>
__global__ void __launch_bounds__(256) active_count_kernel(const uint8_t* __restrict__ s, int64_t m, int64_t n,
                                                           int* __restrict__ colcount) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= n) return;
    const uint8_t* cp = s + (size_t)col * (size_t)m;
    int cnt = 0;
    for (int64_t i0 = (int64_t)lane * 16; i0 < m; i0 += 512) cnt += __popc(load_mask16<VEC16>(cp, i0, m));
    cnt = warp_sum_int(cnt);
    if (lane == 0) colcount[col] = cnt;
}

template <bool VEC16>
__global__ void __launch_bounds__(256) active_fill_kernel(const uint8_t* __restrict__ s, int64_t m, int64_t n,
                                                          const int* __restrict__ colptr, int* __restrict__ yrow,
                                                          int* __restrict__ ycol, int* __restrict__ rowcount) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= n) return;
    const uint8_t* cp = s + (size_t)col * (size_t)m;
    int base = colptr[col];
    for (int64_t w0 = 0; w0 < m; w0 += 512) {
        const int64_t i0 = w0 + (int64_t)lane * 16;
        const unsigned mask = (i0 < m) ? load_mask16<VEC16>(cp, i0, m) : 0u;
        const int mine = __popc(mask);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int pos = base + incl - mine;
        unsigned mm = mask;
        while (mm) {
            const int b = __ffs(mm) - 1; mm &= mm - 1;
            const int r = (int)(i0 + b);
            yrow[pos] = r; ycol[pos] = (int)col;
            atomicAdd(rowcount + r, 1);
            ++pos;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

}  // namespace

// =================================================================== host entry points

void plan_ax(ssn_ctx* c, const double* x, const double* p, const double* q, int64_t m, int64_t n, double* y) {
    SSN_REQUIRE(m > 0 && n > 0 && x && p && q && y, SSN_E_INVALID, "Ax: bad arguments");
    const Tiling t = plan_tiling(c, m, n);
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n);
    PlanArgs a{};
    a.x = x; a.p = p; a.q = q; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups;
    a.rowpart = rowpart; a.colpart = colpart; a.want_sums = 1;
    const dim3 grid(t.chunks, t.groups);
    const bool vec = vec_ok(x, m);
    a.stage = 0;                                            // two FMAs per entry: the register-loading form is the faster one (measured)
    const size_t smem = (size_t)kWarps * t.cpc * sizeof(double) + (a.stage ? kStageBytes : 0);
    if (vec) {
        static bool attr = false;
        if (!attr) { SSN_CUDA(cudaFuncSetAttribute((plan_reduce_kernel<MODE_AX, true, G_INF>), cudaFuncAttributeMaxDynamicSharedMemorySize, kStageSmemMax)); attr = true; }
        SSN_LAUNCH(c, (plan_reduce_kernel<MODE_AX, true, G_INF>), grid, kThreads, smem, a);
    } else SSN_LAUNCH(c, (plan_reduce_kernel<MODE_AX, false, G_INF>), grid, kThreads, smem, a);
    SSN_LAUNCH(c, plan_finish_kernel, cdiv(m + n, 256), 256, 0, rowpart.p, colpart.p, nullptr, t.chunks,
               t.groups, m, n, 0, y, nullptr);
}

void plan_aty(ssn_ctx* c, const double* y, const double* p, const double* q, int64_t m, int64_t n, double* z) {
    SSN_REQUIRE(m > 0 && n > 0 && y && p && q && z, SSN_E_INVALID, "Aty: bad arguments");
    // a store-only kernel has no partial arrays: 8 waves of short blocks (a bare write of the plan at this tiling: 6.27 TB/s
    // with 2 waves, 6.98 TB/s with 8 on a B200; tools/micro/ldst256.cu)
    const Tiling t = plan_tiling(c, m, n, 8);
    const dim3 grid(t.chunks, t.groups);
    if (vec_ok(z, m)) SSN_LAUNCH(c, (aty_kernel<true>), grid, kThreads, 0, y, p, q, m, n, t.cpc, z);
    else              SSN_LAUNCH(c, (aty_kernel<false>), grid, kThreads, 0, y, p, q, m, n, t.cpc, z);
}

void plan_prox_residual(ssn_ctx* c, const double* w, const double* lam, const double* p, const double* q,
                        int64_t m, int64_t n, double tk, const double* gama, double gama_s, double* axp_out,
                        double* prox_out, double* z_out, uint8_t* s_out, double* scal2_dev) {
    SSN_REQUIRE(m > 0 && n > 0 && w && lam && p && q, SSN_E_INVALID, "prox_residual: bad arguments");
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    Buf<double> rowpart, colpart, scalpart(c, (size_t)2 * nblocks);
    if (axp_out) { rowpart.alloc(c, (size_t)t.chunks * m); colpart.alloc(c, (size_t)t.groups * n); }
    PlanArgs a{};
    a.x = w; a.p = p; a.q = q; a.lam = lam; a.gama = gama; a.gama_s = gama_s; a.inv_tk = 1.0 / tk;
    a.m = m; a.n = n; a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups;
    a.rowpart = rowpart.p; a.colpart = colpart.p; a.scalpart = scalpart.p;
    a.prox_out = prox_out; a.z_out = z_out; a.s_out = s_out; a.want_sums = axp_out ? 1 : 0;
    const dim3 grid(t.chunks, t.groups);
    bool vec = vec_ok(w, m) && (!gama || vec_ok(gama, m)) && (!prox_out || vec_ok(prox_out, m)) &&
               (!z_out || vec_ok(z_out, m)) && (!s_out || (reinterpret_cast<uintptr_t>(s_out) & 1u) == 0);
    const int gm = gama ? G_VECTOR : (std::isinf(gama_s) && gama_s > 0 ? G_INF : G_SCALAR);
    // the double buffer pays once a block walks many batches (full-size plan: 111 per block); on the short blocks of a small
    // slab its prologue costs more than it hides (2048 x 16384 slab: 63 us staged against 56 us)
    a.stage = (vec && gm != G_VECTOR && c->plan_stage && t.cpc >= 128) ? 1 : 0;
    const size_t smem = (size_t)kWarps * t.cpc * sizeof(double) + (a.stage ? kStageBytes : 0);
    if (a.stage) {
        static bool attr = false;
        if (!attr) {
            SSN_CUDA(cudaFuncSetAttribute((plan_reduce_kernel<MODE_PROX, true, G_INF>), cudaFuncAttributeMaxDynamicSharedMemorySize, kStageSmemMax));
            SSN_CUDA(cudaFuncSetAttribute((plan_reduce_kernel<MODE_PROX, true, G_SCALAR>), cudaFuncAttributeMaxDynamicSharedMemorySize, kStageSmemMax));
            attr = true;
        }
    }
#define SSN_PROX_LAUNCH(V, G) SSN_LAUNCH(c, (plan_reduce_kernel<MODE_PROX, V, G>), grid, kThreads, smem, a)
    {
    KernelTimer kt(c);
    if (vec) { if (gm == G_INF) SSN_PROX_LAUNCH(true, G_INF); else if (gm == G_SCALAR) SSN_PROX_LAUNCH(true, G_SCALAR); else SSN_PROX_LAUNCH(true, G_VECTOR); }
    else     { if (gm == G_INF) SSN_PROX_LAUNCH(false, G_INF); else if (gm == G_SCALAR) SSN_PROX_LAUNCH(false, G_SCALAR); else SSN_PROX_LAUNCH(false, G_VECTOR); }
    }
#undef SSN_PROX_LAUNCH
    SSN_LAUNCH(c, plan_finish_kernel, axp_out ? cdiv(m + n, 256) : 1, 256, 0, rowpart.p, colpart.p, scalpart.p,
               t.chunks, t.groups, m, n, nblocks, axp_out, scal2_dev);
}

// ---- Class 2 (partial OT): u = [x (m*n) ; y (n) ; z (m)], lk = [lambda (n+m) ; last dual], H = [A  I ; phi' 0]
namespace {
// the slack blocks y, z (N = n + m entries; H'lk there is lk(1:N)) and the last row of H*prox(z): one block
__global__ void __launch_bounds__(1024) pot_slack_kernel(const double* __restrict__ w_s, const double* __restrict__ lam, int N, double inv_tk,
                                                         double* __restrict__ hp, double* __restrict__ prox_s, double* __restrict__ t_out,
                                                         const double* __restrict__ phipart, int nblocks, double* __restrict__ scal3) {
    __shared__ double red[32];
    double n2 = 0.0;
    for (int i = threadIdx.x; i < N; i += 1024) {
        const double z = __dmul_rn(inv_tk, __dsub_rn(w_s[i], lam[i]));                     // 1/tk*(wk - Htlk), slack part
        const bool nonneg = z >= 0.0;
        const double pz = nonneg ? z : 0.0;
        if (hp) hp[i] += pz;                                                                // Ax(prox x) + [prox y ; prox z]
        if (prox_s) prox_s[i] = pz;
        if (t_out) t_out[i] = nonneg ? 1.0 : 0.0;
        n2 = fma(pz, pz, n2);
    }
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 1024) s += phipart[b];
    const double t2 = block_sum(n2, red);
    const double t3 = block_sum(s, red);
    if (threadIdx.x == 0) {
        scal3[0] += t2;                                                                     // ||prox(z)||^2 over x, y and z
        scal3[2] = t3;                                                                      // phi'*prox(z_x)
        if (hp) hp[N] = t3;
    }
}
}  // namespace

// Fused Class 2 residual pieces, one read of w and one of phi (Class2/APD_SsN_Class2.m:124-130, 137-150, 196-217):
//   zk = 1/tk*(wk - [Aty(lk(1:N),p,q) + lk(N+1)*phi ; lk(1:N)]),  s = zk(1:mn) >= 0,  t = zk(mn+1:end) >= 0,
//   Hpzk = [Ax(prox x) + [prox y ; prox z] ; phi'*prox x]  -> hp_out (N+1),  ||prox(zk)||^2, nnz(s)  -> scal3_dev[0..1]
void plan_prox_residual_pot(ssn_ctx* c, const double* w, const double* lam, const double* p, const double* q, int64_t m, int64_t n,
                            double tk, const double* phi, double* hp_out, double* prox_out, uint8_t* s_out, double* t_out,
                            double* scal3_dev) {
    SSN_REQUIRE(m > 0 && n > 0 && w && lam && p && q && phi && scal3_dev, SSN_E_INVALID, "prox_residual_pot: bad arguments");
    SSN_REQUIRE(m + n < (int64_t)1 << 30, SSN_E_TOO_LARGE, "prox_residual_pot: m + n too large");
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    const int64_t mn = m * n, N = m + n;
    Buf<double> rowpart, colpart, scalpart(c, (size_t)2 * nblocks), phipart(c, (size_t)nblocks);
    if (hp_out) { rowpart.alloc(c, (size_t)t.chunks * m); colpart.alloc(c, (size_t)t.groups * n); }
    PlanArgs a{};
    a.x = w; a.p = p; a.q = q; a.lam = lam; a.gama = phi; a.gama_s = 0.0; a.inv_tk = 1.0 / tk;       // the kernel reads lk(N+1) itself
    a.m = m; a.n = n; a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups;
    a.rowpart = rowpart.p; a.colpart = colpart.p; a.scalpart = scalpart.p; a.phipart = phipart.p;
    a.prox_out = prox_out; a.z_out = nullptr; a.s_out = s_out; a.want_sums = hp_out ? 1 : 0;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(w, m) && vec_ok(phi, m) && (!prox_out || vec_ok(prox_out, m)) && (!s_out || (reinterpret_cast<uintptr_t>(s_out) & 1u) == 0);
    {
        KernelTimer kt(c);
        if (vec) SSN_LAUNCH(c, (plan_reduce_kernel<MODE_PROX, true, G_PHI>), grid, kThreads, smem, a);
        else     SSN_LAUNCH(c, (plan_reduce_kernel<MODE_PROX, false, G_PHI>), grid, kThreads, smem, a);
    }
    SSN_LAUNCH(c, plan_finish_kernel, hp_out ? cdiv(N, 256) : 1, 256, 0, rowpart.p, colpart.p, scalpart.p, t.chunks, t.groups, m, n, nblocks,
               hp_out, scal3_dev);
    SSN_LAUNCH(c, pot_slack_kernel, 1, 1024, 0, w + mn, lam, (int)N, 1.0 / tk, hp_out, prox_out ? prox_out + mn : nullptr, t_out,
               phipart.p, nblocks, scal3_dev);
}

// n2_out_dev[t] = ||prox((w - Aty(lamT[t]))/tk)||^2 for t < nt (nt <= kMaxTrials), one read of w
// flag_dev[0] = 1 iff some weight differs from 1 (decides the fp64-lean path of the trials kernel)
const int* plan_nonunit_flag(ssn_ctx* c, const double* p, const double* q, int64_t m, int64_t n, int* flag_dev) {
    SSN_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int), c->stream));
    SSN_LAUNCH(c, unit_weights_kernel, std::min(64, cdiv(m + n, 256)), 256, 0, p, m, q, n, flag_dev);
    return flag_dev;
}

void plan_prox_trials(ssn_ctx* c, const double* w, const double* lamT, int nt, const double* p, const double* q,
                      int64_t m, int64_t n, double tk, const double* gama, double gama_s, double* n2_out_dev,
                      const int* nonunit_dev) {
    SSN_REQUIRE(m > 0 && n > 0 && w && lamT && p && q && n2_out_dev, SSN_E_INVALID, "prox_trials: bad arguments");
    SSN_REQUIRE(nt >= 1 && nt <= kMaxTrials, SSN_E_INVALID, "prox_trials: 1 <= nt <= 8");
    Buf<double> n2_scratch(c, kMaxTrials);
    const int NTk = nt <= 1 ? 1 : (nt <= 2 ? 2 : (nt <= 4 ? 4 : 8));     // compiled batch sizes
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    Buf<double> scalpart(c, (size_t)NTk * nblocks);
    Buf<int> flag;
    if (!nonunit_dev) { flag.alloc(c, 1); nonunit_dev = plan_nonunit_flag(c, p, q, m, n, flag.p); }
    TrialArgs a{};
    a.nonunit = nonunit_dev; a.nt_valid = nt;
    a.w = w; a.p = p; a.q = q; a.lamT = lamT; a.gama = gama; a.gama_s = gama_s; a.inv_tk = 1.0 / tk;
    a.m = m; a.n = n; a.ldl = n + m; a.cols_per_chunk = t.cpc; a.scalpart = scalpart.p;
    const dim3 grid(t.chunks, t.groups);
    const bool vec = vec_ok(w, m) && (!gama || vec_ok(gama, m));
    const int gm = gama ? G_VECTOR : (std::isinf(gama_s) && gama_s > 0 ? G_INF : G_SCALAR);
    const size_t tsmem = sizeof(double) * ((size_t)NTk * kGroupRows + (size_t)(NTk + 1) * t.cpc);
#define SSN_TRIALS_NT(V, G, NT) do { \
        SSN_CUDA(cudaFuncSetAttribute((plan_trials_kernel<V, G, NT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem)); \
        SSN_LAUNCH(c, (plan_trials_kernel<V, G, NT>), grid, kThreads, tsmem, a); } while (0)
#define SSN_TRIALS_G(V, G) do { if (NTk == 1) SSN_TRIALS_NT(V, G, 1); else if (NTk == 2) SSN_TRIALS_NT(V, G, 2); \
                                else if (NTk == 4) SSN_TRIALS_NT(V, G, 4); else SSN_TRIALS_NT(V, G, 8); } while (0)
    {
    KernelTimer kt(c);
    if (vec) { if (gm == G_INF) SSN_TRIALS_G(true, G_INF); else if (gm == G_SCALAR) SSN_TRIALS_G(true, G_SCALAR); else SSN_TRIALS_G(true, G_VECTOR); }
    else     { if (gm == G_INF) SSN_TRIALS_G(false, G_INF); else if (gm == G_SCALAR) SSN_TRIALS_G(false, G_SCALAR); else SSN_TRIALS_G(false, G_VECTOR); }
    }
#undef SSN_TRIALS_G
#undef SSN_TRIALS_NT
    SSN_LAUNCH(c, trials_finish_kernel, 1, 256, 0, scalpart.p, nblocks, NTk, n2_scratch.p);
    SSN_CUDA(cudaMemcpyAsync(n2_out_dev, n2_scratch.p, sizeof(double) * nt, cudaMemcpyDeviceToDevice, c->stream));
}

// Screened trials (screen / compact / eval kernels above): out_dev[t] = ||prox((w - Aty(lam + alpha_t*zeta))/tk)||^2
// for t < nt (nt <= 256, alpha_t = delta^(ll0+t), gama = Inf) and out_dev[nt] = number of entries that
// survived the screen (candidates) -- out of m*n -- which tells the caller how sparse the trial plans are.
// No host round trip before the last kernel is enqueued (the candidate list has a fixed capacity; see below).
void plan_prox_trials_lin(ssn_ctx* c, const double* w, const double* lam, const double* zeta, const double* p, const double* q,
                          int64_t m, int64_t n, double tk, double delta, int ll0, int nt, double* out_dev, const int* nonunit_dev) {
    SSN_REQUIRE(m > 0 && n > 0 && w && lam && zeta && p && q && out_dev, SSN_E_INVALID, "prox_trials_lin: bad arguments");
    SSN_REQUIRE(nt >= 1 && nt <= kMaxLinBatch, SSN_E_INVALID, "prox_trials_lin: 1 <= nt <= 256");
    SSN_REQUIRE(m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31), SSN_E_TOO_LARGE, "prox_trials_lin: m or n >= 2^31");
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    const int words = cdiv(t.cpc, 8);
    const int egrid = c->num_sms * 4;                                       // evaluation grid (fixed: the summation order depends on it)
    Buf<double> scalpart(c, (size_t)kMaxLinTrials * egrid), votepart(c, nblocks), scratch(c, kMaxLinBatch + 1);
    Buf<unsigned> mask(c, (size_t)nblocks * words * kThreads);
    Buf<unsigned long long> off(c, (size_t)nblocks + 1);
    Buf<int> flag;
    if (!nonunit_dev) { flag.alloc(c, 1); nonunit_dev = plan_nonunit_flag(c, p, q, m, n, flag.p); }
    TrialLinArgs a{};
    a.w = w; a.p = p; a.q = q; a.lam = lam; a.zeta = zeta; a.inv_tk = 1.0 / tk; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.words = words; a.mask = mask.p; a.scalpart = scalpart.p; a.votepart = votepart.p; a.nonunit = nonunit_dev;
    a.alpha[0] = std::pow(delta, (double)ll0); a.alpha[1] = std::pow(delta, (double)(ll0 + nt - 1));      // first and last step of the batch
    const dim3 grid(t.chunks, t.groups);
    const bool vec = vec_ok(w, m);
    const size_t tsmem = sizeof(double) * (size_t)3 * t.cpc;
    {
    KernelTimer kt(c);                                   // the plan-wide (HBM-bound) kernel of the batch
    if (vec) SSN_LAUNCH(c, (plan_trials_screen_kernel<true>), grid, kThreads, tsmem, a);
    else     SSN_LAUNCH(c, (plan_trials_screen_kernel<false>), grid, kThreads, tsmem, a);
    }
    SSN_LAUNCH(c, cand_scan_kernel, 1, 1024, 0, votepart.p, nblocks, off.p, scratch.p + nt);
    // The candidate list is sized for 1/16 of the plan (the caller leaves the screened path long before the
    // trial plans are that dense), so the whole batch is enqueued without a host round trip; if the list
    // does not fit, the kernels do nothing and the list pass is repeated with the exact size.
    auto list_pass = [&](int2* list, unsigned long long cap) {
        if (vec) SSN_LAUNCH(c, (cand_compact_kernel<true>), grid, kThreads, 0, a, off.p, list, cap);
        else     SSN_LAUNCH(c, (cand_compact_kernel<false>), grid, kThreads, 0, a, off.p, list, cap);
        for (int t0 = 0; t0 < nt; t0 += kMaxLinTrials) {      // 32 steps per evaluation launch
            const int k = std::min(kMaxLinTrials, nt - t0);
            const int NTk = k <= 1 ? 1 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32));
            for (int i = 0; i < kMaxLinTrials; ++i) a.alpha[i] = std::pow(delta, (double)(ll0 + t0 + std::min(i, k - 1)));
            if (NTk == 1) SSN_LAUNCH(c, (cand_eval_kernel<1>), egrid, kThreads, 0, a, list, off.p + nblocks, cap);
            else if (NTk == 8) SSN_LAUNCH(c, (cand_eval_kernel<8>), egrid, kThreads, 0, a, list, off.p + nblocks, cap);
            else if (NTk == 16) SSN_LAUNCH(c, (cand_eval_kernel<16>), egrid, kThreads, 0, a, list, off.p + nblocks, cap);
            else SSN_LAUNCH(c, (cand_eval_kernel<32>), egrid, kThreads, 0, a, list, off.p + nblocks, cap);
            SSN_LAUNCH(c, trials_lin_finish_kernel, k, 256, 0, scalpart.p, egrid, NTk, 0.25, scratch.p + t0);
        }
    };
    const unsigned long long cap0 = std::max<unsigned long long>(1ull << 16, (unsigned long long)m * (unsigned long long)n / 16);
    {
        Buf<int2> list(c, (size_t)cap0);
        list_pass(list.p, cap0);
    }
    const double total = read_scalar(c, scratch.p + nt);
    if (total > (double)cap0) {
        Buf<int2> list(c, (size_t)total);
        list_pass(list.p, (unsigned long long)total);
    }
    SSN_CUDA(cudaMemcpyAsync(out_dev, scratch.p, sizeof(double) * (nt + 1), cudaMemcpyDeviceToDevice, c->stream));
}

// lamT[t] = lam + delta^(ll0+t)*zeta for t < nt, and f0_out[2t] = ||lamT[t]||^2, f0_out[2t+1] = wlk'*lamT[t]
// (the O(m+n) half of the line-search objective; the row-sharded step calls this on the full dual
// vectors and plan_prox_trials on its slab).
void plan_trial_vectors(ssn_ctx* c, const double* lam, const double* zeta, const double* wlk, int64_t N, double delta,
                        int ll0, int nt, double* lamT, double* f0_out) {
    SSN_REQUIRE(lam && zeta && wlk && lamT && f0_out && nt >= 1 && nt <= kMaxLinBatch && N > 0, SSN_E_INVALID, "trial_vectors: bad arguments");
    const int nb = 64;
    Buf<double> alpha(c, nt), f0part(c, (size_t)nb * nt * 2);
    for (int t = 0; t < nt; ++t) c->h_pin[1024 + t] = std::pow(delta, (double)(ll0 + t));
    SSN_CUDA(cudaMemcpyAsync(alpha.p, c->h_pin + 1024, sizeof(double) * nt, cudaMemcpyHostToDevice, c->stream));
    SSN_LAUNCH(c, trial_vectors_kernel, dim3(nb, nt), 256, 0, N, nt, lam, zeta, wlk, alpha.p, lamT, f0part.p);
    SSN_LAUNCH(c, trial_f0_finish_kernel, 2 * nt, 256, 0, f0part.p, nb, nt, f0_out);
    SSN_CUDA(cudaStreamSynchronize(c->stream));           // h_pin is reused by the next call
}

// Armijo backtracking of Class1/APD_SsN_Class1.m:182-211, kMaxTrials trial steps per read of w:
//   lk_new = lk_old + delta^ll*zeta ; cF_new = bk1/2*||lk_new||^2 - wlk'*lk_new + tk/2*||prox(z)||^2 ;
//   accept the first ll with  !(cF_new > cF_old - nu*delta^ll*ress)  or  ll == ll_max.
void plan_linesearch(ssn_ctx* c, const double* w, const double* lam_old, const double* zeta, const double* wlk,
                     const double* p, const double* q, int64_t m, int64_t n, double tk, double bk1, const double* gama,
                     double gama_s, double nu, double delta, int ll_max, double cF_old, double ress, int batch,
                     double* lam_new, int* ll_out, double* n2_out, double* cF_out, int* passes_out) {
    SSN_REQUIRE(lam_old && zeta && wlk && lam_new && ll_max >= 0, SSN_E_INVALID, "linesearch: bad arguments");
    // batch <= 0: adaptive.  With gama = Inf every pass goes through the screened kernels, whose candidate
    // count says how sparse the trial plans are: while fewer than 10 % of the entries survive the screen
    // a read of w evaluates 64, then 128 backtracking steps (the evaluation of the candidates is a few
    // tens of microseconds per 32 steps), 16 below 25 %, else 8 through the dense kernel.
    const bool adaptive = batch <= 0;
    const bool screened = adaptive && c->ls_screen && gama == nullptr && std::isinf(gama_s) && gama_s > 0;
    if (adaptive) batch = kMaxTrials;
    if (batch > kMaxTrials) batch = kMaxTrials;
    const int cap = screened ? kMaxLinBatch : batch;
    const int64_t N = m + n;
    const int nb = 64;
    Buf<double> lamT(c, (size_t)cap * N), alpha(c, cap), f0part(c, (size_t)nb * cap * 2), res(c, 3 * (size_t)cap + 1);
    Buf<int> flag(c, 1);
    const int* nonunit = plan_nonunit_flag(c, p, q, m, n, flag.p);          // once per line search
    const double slots = std::max(1.0, (double)m * (double)n);
    int ll = 0, passes = 0;
    double dens = 1.0;                                                      // surviving share of the entries, last pass
    // most steps accept the full step (ll = 0): the first read of w evaluates that trial alone -- unless the trial plans
    // of the previous line search of this context were sparse (late phase, where line searches are long): a screened
    // batch of 64 steps costs the same one read of w then, and saves a pass.  Every later read evaluates a batch.
    int first = (screened && c->ls_last_density >= 0.0 && c->ls_last_density <= 0.10) ? std::min(64, c->ls_max_nt) : 1;
    // ... and when the previous line search of this context backtracked far (late phase: consecutive SsN steps accept at similar
    // ll, 100 - 200), the first batch reaches past that ll -- up to 256 steps in the one read of w, rounded to the 32 steps of an
    // evaluation launch -- instead of finding it in a second pass.  The accepted step is the first that passes the Armijo test
    // whatever the batches are.
    if (first > 1 && c->ls_max_nt >= 128 && c->ls_last_ll >= 48) first = std::min(kMaxLinBatch, ((c->ls_last_ll + 32 + 31) / 32) * 32);
    while (true) {
        int want = batch;
        bool lin = screened;
        if (screened && passes > 0) {
            if (dens <= 0.10) want = std::min(c->ls_max_nt, (passes == 1 && first == 1) ? 64 : 128);
            else if (dens <= 0.25) want = std::min(16, c->ls_max_nt);
            else { want = kMaxTrials; lin = false; }
        }
        const int nt = std::min(passes == 0 ? first : want, ll_max - ll + 1);
        double al[kMaxLinBatch];
        for (int t = 0; t < nt; ++t) al[t] = std::pow(delta, (double)(ll + t));
        for (int t = 0; t < nt; ++t) c->h_pin[1024 + t] = al[t];
        SSN_CUDA(cudaMemcpyAsync(alpha.p, c->h_pin + 1024, sizeof(double) * nt, cudaMemcpyHostToDevice, c->stream));
        SSN_LAUNCH(c, trial_vectors_kernel, dim3(nb, nt), 256, 0, N, nt, lam_old, zeta, wlk, alpha.p, lamT.p, f0part.p);
        SSN_LAUNCH(c, trial_f0_finish_kernel, 2 * nt, 256, 0, f0part.p, nb, nt, res.p + cap + 1);
        if (lin) plan_prox_trials_lin(c, w, lam_old, zeta, p, q, m, n, tk, delta, ll, nt, res.p, nonunit);
        else     plan_prox_trials(c, w, lamT, nt, p, q, m, n, tk, gama, gama_s, res.p, nonunit);
        double h[3 * kMaxLinBatch + 1];
        read_back(c, res.p, h, 3 * (size_t)cap + 1);
        if (lin) { dens = h[nt] / slots; c->ls_last_density = dens; }
        ++passes;
        int acc = -1;
        double n2a = 0.0, cFa = 0.0;
        for (int t = 0; t < nt; ++t) {
            const double f0 = bk1 / 2 * h[cap + 1 + 2 * t] - h[cap + 1 + 2 * t + 1];
            const double cF_new = f0 + 0.5 * tk * h[t];
            if (!(cF_new > cF_old - nu * al[t] * ress) || ll + t == ll_max) { acc = t; n2a = h[t]; cFa = cF_new; break; }
        }
        if (acc >= 0) {
            ll += acc;
            SSN_CUDA(cudaMemcpyAsync(lam_new, lamT.p + (size_t)acc * N, sizeof(double) * N, cudaMemcpyDeviceToDevice, c->stream));
            if (n2_out) *n2_out = n2a;
            if (cF_out) *cF_out = cFa;
            break;
        }
        ll += nt;
    }
    c->ls_last_ll = ll;
    if (ll_out) *ll_out = ll;
    if (passes_out) *passes_out = passes;
}

// [xk, lk] = warmup_class1(c,r,l,p,q,gama,0,maxit) -- Class1/warmup_class1.m:18-96 with maxit < inf
// (the call of Class1/APD_SsN_Class1.m:59), device resident, two plan-wide kernels per iteration.
void plan_warmup_class1(ssn_ctx* c, const double* cost, const double* b, const double* p, const double* q, int64_t m,
                        int64_t n, const double* gama, double gama_s, int maxit, double* xk_out, double* lk_out) {
    SSN_REQUIRE(cost && b && p && q && xk_out && lk_out && m > 0 && n > 0 && maxit >= 0, SSN_E_INVALID, "warmup_class1: bad arguments");
    const int64_t N = m + n; const size_t mn = (size_t)m * (size_t)n;
    const Tiling t = plan_tiling(c, m, n);
    Buf<double> vk(c, mn), wk(c, mn), pik(c, mn), lk2(c, mn), dd(c, mn);
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n), rowpart2(c, (size_t)t.chunks * m), colpart2(c, (size_t)t.groups * n);
    Buf<double> axk(c, N), axdd(c, N), av(c, N), y(c, N);
    double* xk = xk_out; double* lk1 = lk_out;
    SSN_CUDA(cudaMemsetAsync(xk, 0, sizeof(double) * mn, c->stream));
    vk.zero(); wk.zero(); pik.zero(); lk2.zero(); axk.zero();
    SSN_CUDA(cudaMemsetAsync(lk1, 0, sizeof(double) * N, c->stream));
    WarmArgs a{};
    a.xk = xk; a.vk = vk; a.wk = wk; a.pik = pik; a.lk2 = lk2; a.dd = dd; a.c = cost; a.p = p; a.q = q; a.b = b;
    a.lk1 = lk1; a.axk = axk; a.y = y; a.gama = gama; a.gama_s = gama_s; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups;
    a.rowpart = rowpart; a.colpart = colpart; a.rowpart2 = rowpart2; a.colpart2 = colpart2;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)2 * kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(xk, m) && vec_ok(cost, m) && (!gama || vec_ok(gama, m)) && vec_ok(vk.p, m) && vec_ok(dd.p, m);
    const int gm = gama ? G_VECTOR : (std::isinf(gama_s) && gama_s > 0 ? G_INF : G_SCALAR);
    const double muf = 0.0;
    double gk = 1.0, bk = 1.0;                                              // warmup_class1.m:27
    for (int k = 1; k <= maxit; ++k) {
        const double ak = bk, bk1 = bk / (1 + ak);                          // :59
        const double gk1 = (gk + muf * ak) / (1 + ak);
        const double etafk = (1 + ak) * gk + muf * ak;
        const double sgk = 1 / bk1, etagk = (1 + ak) * bk;
        const double tt = sgk * ak * ak, sg = 1 + etafk / tt;               // :69
        a.ak = ak; a.bk = bk; a.gk = gk; a.muf = muf; a.etafk = etafk; a.sgk = sgk; a.etagk = etagk; a.tt = tt;
        if (vec) SSN_LAUNCH(c, (warm_kernel<0, true, G_INF>), grid, kThreads, smem / 2, a);
        else     SSN_LAUNCH(c, (warm_kernel<0, false, G_INF>), grid, kThreads, smem / 2, a);
        SSN_LAUNCH(c, plan_finish2_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, nullptr, nullptr, t.chunks, t.groups, m, n, axdd.p, nullptr);
        SSN_LAUNCH(c, invaat_block_kernel, 1, 1024, 0, n, m, axdd.p, p, q, sg, sg, y.p);          // :70
#define SSN_WARM_B(V, G) do { \
        if (k == 1) SSN_CUDA(cudaFuncSetAttribute((warm_kernel<1, V, G>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        SSN_LAUNCH(c, (warm_kernel<1, V, G>), grid, kThreads, smem, a); } while (0)
        if (vec) { if (gm == G_INF) SSN_WARM_B(true, G_INF); else if (gm == G_SCALAR) SSN_WARM_B(true, G_SCALAR); else SSN_WARM_B(true, G_VECTOR); }
        else     { if (gm == G_INF) SSN_WARM_B(false, G_INF); else if (gm == G_SCALAR) SSN_WARM_B(false, G_SCALAR); else SSN_WARM_B(false, G_VECTOR); }
#undef SSN_WARM_B
        SSN_LAUNCH(c, plan_finish2_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, rowpart2.p, colpart2.p, t.chunks, t.groups, m, n, av.p, axk.p);
        SSN_LAUNCH(c, warm_lk1_kernel, cdiv(N, 256), 256, 0, N, ak / bk, av.p, b, lk1);            // :75
        gk = gk1; bk = bk1;                                                 // :77
    }
}

// ONE fused stage of an A-ADMM warm-start iteration (the kernels of plan_warmup_class1) for callers that own
// the loop: the row-sharded driver runs it on its slab and exchanges the column sums between the stages.
// Dual vectors in the slab's own form [column part (n) ; row part of the slab's rows (m)].
//   stage 0 (warmup_class1.m:63-67): reads xk vk wk pik lk2 c, lk1, axk = Ax(xk), b; writes dd; out1 = Ax(dd)
//   stage 1 (:70-75): reads dd xk wk pik lk2, y = invAAt(Ax(dd)); rewrites xk vk wk pik lk2 in place;
//                     out1 = Ax(vk1), out2 = Ax(xk1)
// out1 / out2: column sums over the slab's rows (partial when the plan is sharded), then the rows' own sums.
void plan_warm_stage(ssn_ctx* c, int stage, double* xk, double* vk, double* wk, double* pik, double* lk2, double* dd,
                     const double* cost, const double* p, const double* q, const double* b, const double* lk1, const double* axk,
                     const double* y, int64_t m, int64_t n, const double* gama, double gama_s, double ak, double bk, double gk,
                     double* out1, double* out2) {
    SSN_REQUIRE(xk && vk && wk && pik && lk2 && dd && cost && p && q && b && out1 && m > 0 && n > 0, SSN_E_INVALID, "warm_stage: bad arguments");
    SSN_REQUIRE(stage == 0 ? (lk1 && axk) : (stage == 1 && y && out2), SSN_E_INVALID, "warm_stage: stage 0 needs lk1, axk; stage 1 needs y, out2");
    const int64_t N = m + n;
    const Tiling t = plan_tiling(c, m, n);
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n), rowpart2, colpart2;
    if (stage == 1) { rowpart2.alloc(c, (size_t)t.chunks * m); colpart2.alloc(c, (size_t)t.groups * n); }
    const double muf = 0.0;                                                 // warmup_class1.m:27
    const double bk1 = bk / (1 + ak);
    WarmArgs a{};
    a.xk = xk; a.vk = vk; a.wk = wk; a.pik = pik; a.lk2 = lk2; a.dd = dd; a.c = cost; a.p = p; a.q = q; a.b = b;
    a.lk1 = lk1; a.axk = axk; a.y = y; a.gama = gama; a.gama_s = gama_s; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups;
    a.rowpart = rowpart; a.colpart = colpart; a.rowpart2 = rowpart2; a.colpart2 = colpart2;
    a.ak = ak; a.bk = bk; a.gk = gk; a.muf = muf; a.etafk = (1 + ak) * gk + muf * ak; a.sgk = 1 / bk1; a.etagk = (1 + ak) * bk;
    a.tt = a.sgk * ak * ak;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)2 * kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(xk, m) && vec_ok(cost, m) && (!gama || vec_ok(gama, m)) && vec_ok(vk, m) && vec_ok(dd, m) &&
                     vec_ok(wk, m) && vec_ok(pik, m) && vec_ok(lk2, m);
    const int gm = gama ? G_VECTOR : (std::isinf(gama_s) && gama_s > 0 ? G_INF : G_SCALAR);
    if (stage == 0) {
        if (vec) SSN_LAUNCH(c, (warm_kernel<0, true, G_INF>), grid, kThreads, smem / 2, a);
        else     SSN_LAUNCH(c, (warm_kernel<0, false, G_INF>), grid, kThreads, smem / 2, a);
        SSN_LAUNCH(c, plan_finish2_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, nullptr, nullptr, t.chunks, t.groups, m, n, out1, nullptr);
        return;
    }
#define SSN_WARM_B(V, G) do { \
        SSN_CUDA(cudaFuncSetAttribute((warm_kernel<1, V, G>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        SSN_LAUNCH(c, (warm_kernel<1, V, G>), grid, kThreads, smem, a); } while (0)
    if (vec) { if (gm == G_INF) SSN_WARM_B(true, G_INF); else if (gm == G_SCALAR) SSN_WARM_B(true, G_SCALAR); else SSN_WARM_B(true, G_VECTOR); }
    else     { if (gm == G_INF) SSN_WARM_B(false, G_INF); else if (gm == G_SCALAR) SSN_WARM_B(false, G_SCALAR); else SSN_WARM_B(false, G_VECTOR); }
#undef SSN_WARM_B
    SSN_LAUNCH(c, plan_finish2_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, rowpart2.p, colpart2.p, t.chunks, t.groups, m, n, out1, out2);
}

// wk = -c + bk*(xk+ak*vk)/ak^2 and axk = Ax(xk)      (Class1/APD_SsN_Class1.m:125-126)
void plan_apd_begin(ssn_ctx* c, const double* cost, const double* xk, const double* vk, const double* p, const double* q,
                    int64_t m, int64_t n, double ak, double bk, double* wk_out, double* axk_out) {
    SSN_REQUIRE(cost && xk && vk && p && q && wk_out && axk_out && m > 0 && n > 0, SSN_E_INVALID, "apd_begin: bad arguments");
    const Tiling t = plan_tiling(c, m, n);
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n);
    ApdArgs a{};
    a.c = cost; a.xk = xk; a.vk = vk; a.wk_out = wk_out; a.p = p; a.q = q; a.ak = ak; a.bk = bk; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups; a.rowpart = rowpart; a.colpart = colpart;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(cost, m) && vec_ok(xk, m) && vec_ok(vk, m) && vec_ok(wk_out, m);
    if (vec) SSN_LAUNCH(c, (apd_kernel<0, true, G_INF>), grid, kThreads, smem, a);
    else     SSN_LAUNCH(c, (apd_kernel<0, false, G_INF>), grid, kThreads, smem, a);
    SSN_LAUNCH(c, plan_finish_kernel, cdiv(m + n, 256), 256, 0, rowpart.p, colpart.p, nullptr, t.chunks, t.groups, m, n, 0, axk_out, nullptr);
}

// xk1 = prox((wk-Aty(lam))/tk), vk1 = xk1+(xk1-xk)/ak, axk1 = Ax(xk1), scal2 = {c'*xk1, ||xk1-prox(xk1-c-Aty(lam))||^2}
// (Class1/APD_SsN_Class1.m:239-254)
void plan_apd_end(ssn_ctx* c, const double* cost, const double* wk, const double* xk, const double* lam, const double* p,
                  const double* q, int64_t m, int64_t n, double tk, double ak, const double* gama, double gama_s, double* xk1,
                  double* vk1, double* axk1_out, double* scal2_dev) {
    SSN_REQUIRE(cost && wk && xk && lam && p && q && xk1 && vk1 && axk1_out && scal2_dev && m > 0 && n > 0, SSN_E_INVALID, "apd_end: bad arguments");
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n), scalpart(c, (size_t)2 * nblocks);
    ApdArgs a{};
    a.c = cost; a.xk = xk; a.wk_in = wk; a.xk1 = xk1; a.vk1 = vk1; a.p = p; a.q = q; a.lam = lam; a.gama = gama; a.gama_s = gama_s;
    a.ak = ak; a.inv_tk = 1.0 / tk; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups; a.rowpart = rowpart; a.colpart = colpart; a.scalpart = scalpart;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(cost, m) && vec_ok(xk, m) && vec_ok(wk, m) && vec_ok(xk1, m) && vec_ok(vk1, m) && (!gama || vec_ok(gama, m));
    const int gm = gama ? G_VECTOR : (std::isinf(gama_s) && gama_s > 0 ? G_INF : G_SCALAR);
#define SSN_APD_E(V, G) SSN_LAUNCH(c, (apd_kernel<1, V, G>), grid, kThreads, smem, a)
    if (vec) { if (gm == G_INF) SSN_APD_E(true, G_INF); else if (gm == G_SCALAR) SSN_APD_E(true, G_SCALAR); else SSN_APD_E(true, G_VECTOR); }
    else     { if (gm == G_INF) SSN_APD_E(false, G_INF); else if (gm == G_SCALAR) SSN_APD_E(false, G_SCALAR); else SSN_APD_E(false, G_VECTOR); }
#undef SSN_APD_E
    SSN_LAUNCH(c, plan_finish_kernel, cdiv(m + n, 256), 256, 0, rowpart.p, colpart.p, scalpart.p, t.chunks, t.groups, m, n, nblocks,
               axk1_out, scal2_dev);
}

// [uk, lk] = warmup_class2(c,r,l,p,q,mu,phi,0,maxit) -- Class2/warmup_class2.m:18-108 (the call of Class2/APD_SsN_Class2.m:50),
// device resident: per A-ADMM iteration two plan-wide kernels over the x block (the PHI variants of warm_kernel: phi rides
// along as one more streamed array, the rank-2 terms and lk(N+1)*phi are formed on the fly), one block for the slack blocks
// and the (N+1)-vectors, invHHt in one block.  b = [r ; l ; mu] (N+1); uk_out has m*n + N entries, lk_out N+1.
void plan_warmup_class2(ssn_ctx* c, const double* cost, const double* b, const double* p, const double* q, int64_t m, int64_t n,
                        const double* phi, int maxit, double* uk_out, double* lk_out) {
    SSN_REQUIRE(cost && b && p && q && phi && uk_out && lk_out && m > 0 && n > 0 && maxit >= 0, SSN_E_INVALID, "warmup_class2: bad arguments");
    SSN_REQUIRE(m + n < (int64_t)1 << 30, SSN_E_TOO_LARGE, "warmup_class2: m + n too large");
    const int64_t N = m + n; const size_t mn = (size_t)m * (size_t)n, L = mn + (size_t)N;
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    Buf<double> vk(c, L), wk(c, L), pik(c, L), lkB(c, L), dd(c, L);
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n), rowpart2(c, (size_t)t.chunks * m), colpart2(c, (size_t)t.groups * n);
    Buf<double> huk(c, N + 1), hdd(c, N + 1), ff(c, N + 1), av(c, N), au(c, N), lphi(c, N), phipart(c, (size_t)2 * nblocks);
    double* uk = uk_out; double* lkA = lk_out;
    SSN_CUDA(cudaMemsetAsync(uk, 0, sizeof(double) * L, c->stream));
    vk.zero(); wk.zero(); pik.zero(); lkB.zero(); huk.zero();
    SSN_CUDA(cudaMemsetAsync(lkA, 0, sizeof(double) * (N + 1), c->stream));
    plan_ax(c, phi, p, q, m, n, lphi);                                      // invHHt.m:8: l = Ax(phi), norm(phi)^2
    const double nphi2 = dev_dot(c, phi, phi, (int64_t)mn);
    WarmArgs a{};
    a.xk = uk; a.vk = vk; a.wk = wk; a.pik = pik; a.lk2 = lkB; a.dd = dd; a.c = cost; a.p = p; a.q = q; a.b = b;
    a.lk1 = lkA; a.axk = huk; a.y = ff; a.gama = nullptr; a.gama_s = INFINITY; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups;
    a.rowpart = rowpart; a.colpart = colpart; a.rowpart2 = rowpart2; a.colpart2 = colpart2; a.phi = phi; a.phipart = phipart;
    PotWarmSlack sl{};
    sl.us = uk + mn; sl.vs = vk.p + mn; sl.ws = wk.p + mn; sl.pis = pik.p + mn; sl.lkBs = lkB.p + mn; sl.dds = dd.p + mn;
    sl.b = b; sl.lkA = lkA; sl.huk = huk; sl.hdd = hdd; sl.ff = ff; sl.av = av; sl.au = au; sl.phipart = phipart; sl.nblocks = nblocks; sl.N = (int)N;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)2 * kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(uk, m) && vec_ok(cost, m) && vec_ok(phi, m) && vec_ok(vk.p, m) && vec_ok(dd.p, m);
    const double muf = 0.0;
    double gk = 1.0, bk = 1.0;                                              // warmup_class2.m:26
    for (int k = 1; k <= maxit; ++k) {
        const double ak = bk, bk1 = bk / (1 + ak);                          // :59-62
        const double gk1 = (gk + muf * ak) / (1 + ak);
        const double etafk = (1 + ak) * gk + muf * ak;
        const double sgk = 1 / bk1, etagk = (1 + ak) * bk;
        const double tt = sgk * ak * ak, sg = 1 + etafk / tt;               // :71
        a.ak = ak; a.bk = bk; a.gk = gk; a.muf = muf; a.etafk = etafk; a.sgk = sgk; a.etagk = etagk; a.tt = tt;
        sl.ak = ak; sl.bk = bk; sl.gk = gk; sl.muf = muf; sl.etafk = etafk; sl.sgk = sgk; sl.etagk = etagk; sl.tt = tt;
        if (vec) SSN_LAUNCH(c, (warm_kernel<0, true, G_INF, true>), grid, kThreads, smem / 2, a);
        else     SSN_LAUNCH(c, (warm_kernel<0, false, G_INF, true>), grid, kThreads, smem / 2, a);
        SSN_LAUNCH(c, plan_finish2_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, nullptr, nullptr, t.chunks, t.groups, m, n, hdd.p, nullptr);
        SSN_LAUNCH(c, pot_warm_slack_kernel<0>, 1, 1024, 0, sl);
        SSN_LAUNCH(c, invhht_block_kernel, 1, 1024, 0, n, m, hdd.p, p, q, sg, lphi.p, nphi2, ff.p);   // :73
        if (vec) {
            if (k == 1) SSN_CUDA(cudaFuncSetAttribute((warm_kernel<1, true, G_INF, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SSN_LAUNCH(c, (warm_kernel<1, true, G_INF, true>), grid, kThreads, smem, a);
        } else {
            if (k == 1) SSN_CUDA(cudaFuncSetAttribute((warm_kernel<1, false, G_INF, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SSN_LAUNCH(c, (warm_kernel<1, false, G_INF, true>), grid, kThreads, smem, a);
        }
        SSN_LAUNCH(c, plan_finish2_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, rowpart2.p, colpart2.p, t.chunks, t.groups, m, n, av.p, au.p);
        SSN_LAUNCH(c, pot_warm_slack_kernel<1>, 1, 1024, 0, sl);
        gk = gk1; bk = bk1;                                                 // :84
    }
}

// Class2/APD_SsN_Class2.m:121-122 in one pass over the x block: wk = -wc + bk*(uk+ak*vk)/ak^2 (m*n + N),
// huk = H*uk = [Ax(xk) + [yk;zk] ; phi'*xk] and wlk = bk1*(lk - 1/bk*(huk - b)) - b (N+1 each)
void plan_apd_begin_pot(ssn_ctx* c, const double* cost, const double* uk, const double* vk, const double* p, const double* q, int64_t m,
                        int64_t n, const double* phi, const double* b, const double* lk, double ak, double bk, double bk1, double* wk_out,
                        double* huk_out, double* wlk_out) {
    SSN_REQUIRE(cost && uk && vk && p && q && phi && b && lk && wk_out && huk_out && wlk_out && m > 0 && n > 0, SSN_E_INVALID, "apd_begin_pot: bad arguments");
    const int64_t N = m + n; const size_t mn = (size_t)m * (size_t)n;
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n), phipart(c, (size_t)nblocks);
    ApdArgs a{};
    a.c = cost; a.xk = uk; a.vk = vk; a.wk_out = wk_out; a.p = p; a.q = q; a.ak = ak; a.bk = bk; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups; a.rowpart = rowpart; a.colpart = colpart;
    a.phi = phi; a.phipart = phipart;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(cost, m) && vec_ok(uk, m) && vec_ok(vk, m) && vec_ok(wk_out, m) && vec_ok(phi, m);
    if (vec) SSN_LAUNCH(c, (apd_kernel<0, true, G_INF, true>), grid, kThreads, smem, a);
    else     SSN_LAUNCH(c, (apd_kernel<0, false, G_INF, true>), grid, kThreads, smem, a);
    SSN_LAUNCH(c, plan_finish_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, nullptr, t.chunks, t.groups, m, n, 0, huk_out, nullptr);
    PotApdSlack sl{};
    sl.us = uk + mn; sl.vs = vk + mn; sl.ws_out = wk_out + mn; sl.lk = lk; sl.b = b; sl.axk = huk_out; sl.wlk = wlk_out;
    sl.phipart = phipart; sl.nblocks = nblocks; sl.ak = ak; sl.bk = bk; sl.bk1 = bk1; sl.N = (int)N; sl.n = (int)n;
    SSN_LAUNCH(c, pot_apd_slack_kernel<0>, 1, 1024, 0, sl);
}

// Class2/APD_SsN_Class2.m:231-238 in one pass over the x block: uk1 = prox(zk) at the duals lk (N+1), vk1 = uk1+(uk1-uk)/ak,
// huk1 = H*uk1 and scal5 = { c'*xk1, ||xk1 - max(xk1-c-(Aty(lk)+lk(N+1)*phi),0)||^2, the same for y and z, ||huk1 - b||^2 }
void plan_apd_end_pot(ssn_ctx* c, const double* cost, const double* wk, const double* uk, const double* lk, const double* p, const double* q,
                      int64_t m, int64_t n, const double* phi, const double* b, double tk, double ak, double* uk1, double* vk1,
                      double* huk1_out, double* scal5_dev) {
    SSN_REQUIRE(cost && wk && uk && lk && p && q && phi && b && uk1 && vk1 && huk1_out && scal5_dev && m > 0 && n > 0, SSN_E_INVALID,
                "apd_end_pot: bad arguments");
    const int64_t N = m + n; const size_t mn = (size_t)m * (size_t)n;
    const Tiling t = plan_tiling(c, m, n);
    const int nblocks = t.chunks * t.groups;
    Buf<double> rowpart(c, (size_t)t.chunks * m), colpart(c, (size_t)t.groups * n), scalpart(c, (size_t)2 * nblocks), phipart(c, (size_t)nblocks);
    ApdArgs a{};
    a.c = cost; a.xk = uk; a.wk_in = wk; a.xk1 = uk1; a.vk1 = vk1; a.p = p; a.q = q; a.lam = lk; a.gama = nullptr; a.gama_s = INFINITY;
    a.ak = ak; a.inv_tk = 1.0 / tk; a.m = m; a.n = n;
    a.cols_per_chunk = t.cpc; a.num_chunks = t.chunks; a.num_groups = t.groups; a.rowpart = rowpart; a.colpart = colpart; a.scalpart = scalpart;
    a.phi = phi; a.phipart = phipart;
    const dim3 grid(t.chunks, t.groups);
    const size_t smem = (size_t)kWarps * t.cpc * sizeof(double);
    const bool vec = vec_ok(cost, m) && vec_ok(uk, m) && vec_ok(wk, m) && vec_ok(uk1, m) && vec_ok(vk1, m) && vec_ok(phi, m);
    if (vec) SSN_LAUNCH(c, (apd_kernel<1, true, G_INF, true>), grid, kThreads, smem, a);
    else     SSN_LAUNCH(c, (apd_kernel<1, false, G_INF, true>), grid, kThreads, smem, a);
    SSN_LAUNCH(c, plan_finish_kernel, cdiv(N, 256), 256, 0, rowpart.p, colpart.p, scalpart.p, t.chunks, t.groups, m, n, nblocks, huk1_out, scal5_dev);
    PotApdSlack sl{};
    sl.us = uk + mn; sl.ws_in = wk + mn; sl.us1 = uk1 + mn; sl.vs1 = vk1 + mn; sl.lk = lk; sl.b = b; sl.axk = huk1_out; sl.scal = scal5_dev;
    sl.phipart = phipart; sl.nblocks = nblocks; sl.ak = ak; sl.inv_tk = 1.0 / tk; sl.N = (int)N; sl.n = (int)n;
    SSN_LAUNCH(c, pot_apd_slack_kernel<1>, 1, 1024, 0, sl);
}

// Y = sparse(reshape(s,m,n)) as two sorted coordinate lists (ASAt.m:15):
//   CSC: colptr[n+1], yrow[E] (rows ascending inside a column), ycol[E]
//   row counts rowcount[m] (for the CSR built by the caller with a stable sort by row)
int64_t plan_active_set(ssn_ctx* c, const uint8_t* s, int64_t m, int64_t n, Buf<int>& colptr, Buf<int>& yrow,
                        Buf<int>& ycol, Buf<int>& rowcount) {
    SSN_REQUIRE(m > 0 && n > 0 && s, SSN_E_INVALID, "active set: bad arguments");
    SSN_REQUIRE(m < (int64_t)1 << 31 && n < (int64_t)1 << 31, SSN_E_TOO_LARGE, "active set: m or n >= 2^31");
    Buf<int> colcount(c, n);
    colptr.alloc(c, n + 1);
    rowcount.alloc(c, m); rowcount.zero();
    const bool v16 = (m % 16 == 0) && ((reinterpret_cast<uintptr_t>(s) & 15u) == 0);
    const int grid = cdiv(n, 8);
    if (v16) SSN_LAUNCH(c, (active_count_kernel<true>), grid, 256, 0, s, m, n, colcount.p);
    else     SSN_LAUNCH(c, (active_count_kernel<false>), grid, 256, 0, s, m, n, colcount.p);
    const int64_t E = scan_counts_to_ptr(c, colcount, colptr, n);
    SSN_REQUIRE(E < ((int64_t)1 << 30), SSN_E_TOO_LARGE, "active set: nnz(s) >= 2^30");
    yrow.alloc(c, E); ycol.alloc(c, E);
    if (E > 0) {
        if (v16) SSN_LAUNCH(c, (active_fill_kernel<true>), grid, 256, 0, s, m, n, colptr.p, yrow.p, ycol.p, rowcount.p);
        else     SSN_LAUNCH(c, (active_fill_kernel<false>), grid, 256, 0, s, m, n, colptr.p, yrow.p, ycol.p, rowcount.p);
    }
    return E;
}

}  // namespace ssn
