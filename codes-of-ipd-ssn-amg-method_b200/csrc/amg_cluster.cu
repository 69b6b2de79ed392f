// amg_cluster.cu -- Class_AMG's solve loop (AMG/Class_AMG.m:89-107 with AMG/MG_Wcycle.m / MG_Vcycle.m) inside ONE
// thread-block cluster, every level vector in DISTRIBUTED SHARED MEMORY.
//
// A late-phase SsN system (N ~ 3e4 rows, ~1e5 nonzeros per level) makes the cycle a chain of ~115 dependent passes of a
// few microseconds of work each: what a pass costs is the latency of its gathers and of the barrier behind it.  The
// grid-wide kernel (amg_solve.cu) pays an L2 round trip per gather, an L2 round trip for the stores before the barrier
// and a grid barrier; the first cluster kernel replaced the barrier only.  Here
//   * the rows of every level are split over the 16 CTAs of one cluster and the level's vectors (e, r, g, the Jacobi
//     ping-pong copy, the solution x) live in the owners' shared memory: a gather is ld.shared::cluster (~200 cycles
//     instead of ~700), a store is a local st.shared, and the release of barrier.cluster has no global store to drain;
//   * each matrix gets, once per launch, a table `loc` that holds for every entry the owner CTA and the byte offset of
//     its column inside the owner's slice, so a gather is mapa + ld with no division;
//   * the block Gauss-Seidel smoother of the bigraph level 1 (Class_AMG.m:48-59) runs in its two-half-sweep form
//     e_f = c + invV*(r_f - U*e_c - c*Axi_f), e_c = invT*(r_c - U'*e_f), which is the reference's
//     e + c*xi + R*(g - c*Axi) with R = [invV 0; -invT*U'*invV invT] multiplied out (V, T diagonal): a third of the
//     gathers of residual + coupled update;
//   * what a sweep reads again and again never leaves the SM: row pointers, 1/diag and A*ones of a CTA's rows are staged
//     in its shared memory, and inside a smoothing loop the first entries of a thread's rows (gather location + value)
//     stay in REGISTERS across the sweeps -- every barrier.cluster acquire invalidates the L1, so anything fetched
//     from global memory inside a pass costs an L2 round trip on the critical path;
//   * the visiting order of the cycle is a host-made op list, reductions ride on the barrier through remote stores.
// Levels below `kd` are applied as the dense cycle operator B_kd built by amg_solve.cu (build_dense_tail).
// The kernel text also runs under the host emulation of tests/emu (SSN_EMU: a cluster of 16 x 64 host threads).
#include "amg.cuh"

namespace ssn {

namespace {

#ifdef SSN_EMU
constexpr int kZT = 64;                         // threads per CTA
typedef uintptr_t zaddr;                        // address of a shared-memory location of some CTA of the cluster
#else
constexpr int kZT = 512;                        // 128 registers per thread: two row slots of a smoothing loop stay in registers
typedef uint32_t zaddr;
#endif
constexpr int kZMaxL = 10;                      // explicit levels + the dense leaf
constexpr int kZCta = 16;
constexpr int kZSlots = 2;                      // rows per thread kept in registers by the smoothing loops
constexpr int kZProgMax = 1024;
constexpr int kZXs = 2048;                      // largest dense leaf

enum { Z_PRE = 0, Z_RESTRICT = 1, Z_LEAF = 2, Z_PROLONG = 3, Z_POST = 4, Z_PCG = 5 };
enum { ZV_E = 0, ZV_R = 1, ZV_G = 2, ZV_ALT = 3 };          // vector slots of a level (x: ZLevel::xslot, level 0 only)

struct ZMat { const int* rp; const int* ci; const double* cv; uint32_t* loc; };
struct ZLevel {
    int N, Nf;                                  // Nf: rows of the first segment (bigraph level: fnode; else N)
    int rpf, rpc;                               // rows per CTA of the two segments
    int so[5];                                  // byte offset of vector slot s in a CTA's shared memory (slot 0 = E first)
    int xslot, bigph;
    int hb, hcap, hn, hoff;                     // halo of the gathered slots: byte offset of the halo area inside such a slot, its capacity
                                                // (entries), the entries in use (set by the kernel, 0: gathers go to the owners), offset of
                                                // the source list (uint32 per entry) in shared memory
    int poff, rsbytes;                          // per-row block: int rs[nl + 2] (row bounds of A, both segments), dinv[nl], Axi[nl]
    int ltA, ltG, ltP, ltT;                     // log2(lanes per row): A (all rows), A (one segment), Pu, Td
    ZMat A, Pu, Td;                             // A_k ; Pro_{k+1} (rows of level k) ; Pro_k' (rows of level k)
    const double* dinv; const double* Axi; double xx; const double* B;
};
struct ZArgs {
    ZLevel lv[kZMaxL];
    int kd, smoth, isnsp, maxit;
    const double* b; double* x; double retol;
    double* relk; double* rho; int* it_out;     // it_out[2]: 0 ok, 1 the level-1 matrix is not [diag U; U' diag] (kernel did nothing)
    const int* prog; int nprog;
    int noreg;                                  // development aid (env SSN_DSM_NOREG=1): the smoothing loops re-read their rows every sweep
    // two-level method (twogrid_bigph.m:98-99): the leaf level kd is solved by PCG(A_kd, r, zero guess, Jacobi) instead of a
    // dense operator; its per-row block then holds diag(A_kd) (not 1/diag) and the level has five vector slots
    int leaf_pcg, pcg_maxit; double pcg_tol2;
    int tail_off, tail_cap;                     // PCG leaf: shared-memory area (byte offset, entries of 12 bytes) for the row entries past the registers
};

constexpr int kOffSlots = 1024, kOffSumR = 1536, kOffDot = 1616, kOffCur = 1696, kOffOuter = 1744, kOffTailCnt = 1792, kOffLv = 2048, kOffProg = 5120,
              kOffXs = 9216, kOffVec = kOffXs + kZXs * 8;
static_assert(sizeof(ZLevel) * kZMaxL <= kOffProg - kOffLv, "level table does not fit its slot");

struct ZTeam {
    int rank, ncta, flip;
    unsigned char* dsm;                         // this CTA's dynamic shared memory
    zaddr base;                                 // the same, as a shared-window address
};

#ifdef SSN_EMU
__device__ __forceinline__ zaddr z_local(const void* p) { return (zaddr)p; }
__device__ __forceinline__ zaddr z_map(zaddr a, int rank) { return (zaddr)emu::cluster_smem[rank] + (a - (zaddr)emu::cluster_smem[blockIdx.x]); }
__device__ __forceinline__ double z_ld(zaddr a) { return *reinterpret_cast<const volatile double*>(a); }
__device__ __forceinline__ void z_st(zaddr a, double v) { *reinterpret_cast<volatile double*>(a) = v; }
__device__ __forceinline__ void z_barrier() { emu::cluster_bar->arrive_and_wait(); }
__device__ __forceinline__ int z_rank() { return blockIdx.x; }
__device__ __forceinline__ int z_ncta() { return gridDim.x; }
#else
__device__ __forceinline__ zaddr z_local(const void* p) { return (zaddr)__cvta_generic_to_shared(p); }
__device__ __forceinline__ zaddr z_map(zaddr a, int rank) { zaddr r; asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
// volatile, no memory clobber: ordered against the barriers (volatile asm with a clobber), free to overlap with the
// plain loads / stores of a pass, none of which touches a location another CTA gathers in the same pass
__device__ __forceinline__ double z_ld(zaddr a) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void z_st(zaddr a, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
// release / acquire at cluster scope: shared-memory writes of every CTA before the barrier are visible to all after it
__device__ __forceinline__ void z_barrier() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ int z_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ int z_ncta() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return (int)r; }
#endif

// global row of local row l of this CTA (-1: past the end of a segment)
__device__ __forceinline__ int z_row(const ZLevel& L, int rank, int l) {
    if (l < L.rpf) { const int r = rank * L.rpf + l; return r < L.Nf ? r : -1; }
    const int r = L.Nf + rank * L.rpc + (l - L.rpf);
    return (l < L.rpf + L.rpc && r < L.N) ? r : -1;
}
// owner CTA (bits 24..), second-segment flag of the bigraph level (bit 23) and byte offset inside the owner's vector slot
// (bits 0..22) of element j of a vector of level L
constexpr uint32_t kZSeg = 1u << 23, kZOffMask = kZSeg - 1u;
__device__ __forceinline__ uint32_t z_loc(const ZLevel& L, int j) {
    int owner, l;
    if (j < L.Nf) { owner = j / L.rpf; l = j - owner * L.rpf; }
    else { const int jj = j - L.Nf; owner = jj / L.rpc; l = L.rpf + jj - owner * L.rpc; }
    return ((uint32_t)owner << 24) | ((L.bigph && j >= L.Nf) ? kZSeg : 0u) | (uint32_t)(l * 8);
}
__device__ __forceinline__ double* z_vec(const ZTeam& G, const ZLevel& L, int slot) { return reinterpret_cast<double*>(G.dsm + L.so[slot]); }
__device__ __forceinline__ int z_vb(const ZLevel& L, int slot) { return L.so[slot]; }
// One element of a level vector: mapa + ld.shared::cluster.  (Measured on a B200, tools/barrier_bench.py: with 16 random
// owner CTAs per warp instruction a 512-thread CTA completes only ~0.3 such loads per cycle -- against ~1 from L2 and ~10
// from its own shared memory -- so what makes this kernel work is the locality of the OT graphs in the natural order:
// most lanes of a warp instruction hit one or two owners.)
__device__ __forceinline__ double z_gather(const ZTeam& G, int vb, uint32_t lc) {
    return z_ld(z_map(G.base + (zaddr)(vb + (int)(lc & kZOffMask)), (int)(lc >> 24)));
}

// cluster-wide sums of two per-thread values: ONE cluster barrier; fixed order, identical in every thread of the cluster
__device__ __forceinline__ void z_sum2(ZTeam& G, double& a, double& b) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = kZT / 32;
    a = warp_sum(a); b = warp_sum(b);
    double* sm = reinterpret_cast<double*>(G.dsm) + (G.flip & 1) * 64;
    if (lane == 0) { sm[w] = a; sm[32 + w] = b; }
    __syncthreads();
    double* sl = reinterpret_cast<double*>(G.dsm + kOffSlots) + (G.flip & 1) * (2 * kZCta);
    if (w == 0) {
        double ta = (lane < NW) ? sm[lane] : 0.0, tb = (lane < NW) ? sm[32 + lane] : 0.0;
        ta = warp_sum(ta); tb = warp_sum(tb);
        if (lane < G.ncta) {
            const zaddr remote = z_map(z_local(sl + 2 * G.rank), lane);
            z_st(remote, ta); z_st(remote + 8, tb);
        }
    }
    z_barrier();
    double s0 = 0.0, s1 = 0.0;
    for (int r = 0; r < G.ncta; ++r) { s0 += sl[2 * r]; s1 += sl[2 * r + 1]; }
    a = s0; b = s1;
    ++G.flip;
}

__device__ __forceinline__ const int* z_rs(const ZTeam& G, const ZLevel& L) { return reinterpret_cast<const int*>(G.dsm + L.poff); }
__device__ __forceinline__ const double* z_dinv(const ZTeam& G, const ZLevel& L) { return reinterpret_cast<const double*>(G.dsm + L.poff + L.rsbytes); }
__device__ __forceinline__ const double* z_axi(const ZTeam& G, const ZLevel& L) { return z_dinv(G, L) + (L.rpf + L.rpc); }
// entries [e0, e1) of local row l in the staged row bounds (the two segments are stored back to back, one sentinel each)
__device__ __forceinline__ void z_bounds(const int* rs, const ZLevel& L, int l, int& e0, int& e1) {
    const int i = l + (l >= L.rpf ? 1 : 0);
    e0 = rs[i]; e1 = rs[i + 1];
}
__device__ __forceinline__ bool z_keep(uint32_t lc, int filt, int /*segb*/) {
    const bool second = (lc & kZSeg) != 0u;
    return filt == 0 || (second == (filt == 1));
}

#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
__device__ unsigned long long g_zdbg[256];
#define ZT0() const long long zt0__ = clock64()
#define ZTACC(slot) do { if (G.rank == 0 && threadIdx.x == 0) { g_zdbg[6 * 16 + (slot)] += (unsigned long long)(clock64() - zt0__); g_zdbg[128 + 6 * 16 + (slot)] += 1ull; } } while (0)
#else
#define ZT0() do {} while (0)
#define ZTACC(slot) do {} while (0)
#endif
// the same for ONE value (most reductions of the cycle).  Stage 1: warp sums into shared memory, every thread of warp 0
// adds the NW partials itself; stage 2: lanes 0..15 of warp 0 put the CTA's sum into slot [rank] of every CTA, ONE
// cluster barrier, every thread adds the 16 slots.  (Sending the warp partials straight to every CTA -- no stage 1 --
// was measured slower: 256 remote 8-byte stores per CTA cost more than the __syncthreads they save.)
// The two halves of z_sum1: post = everything up to and including the barrier, read = the sum of the 16 slots.  (Issuing the
// gathers of the next sweep between the two -- they need the barrier, not the sum -- was measured: slower, the slot reads
// then queue behind the remote loads in the load / store unit.)
__device__ __forceinline__ void z_sum1_post(ZTeam& G, double a) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = kZT / 32;
    ZT0();
    a = warp_sum(a);
    double* sm = reinterpret_cast<double*>(G.dsm) + (G.flip & 1) * 64;
    if (lane == 0) sm[w] = a;
    __syncthreads();
    ZTACC(0);                                                           // debug build: thread 0 waits for the CTA's slowest thread
    double* sl = reinterpret_cast<double*>(G.dsm + kOffSlots) + (G.flip & 1) * (2 * kZCta);
    if (w == 0 && lane < G.ncta) {
        double ta = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) ta += sm[i];
        z_st(z_map(z_local(sl + 2 * G.rank), lane), ta);
    }
    ZTACC(1);
    z_barrier();
    ZTACC(2);                                                           // ... and for the cluster's slowest CTA
}
__device__ __forceinline__ double z_sum1_read(ZTeam& G) {
    const double* sl = reinterpret_cast<const double*>(G.dsm + kOffSlots) + (G.flip & 1) * (2 * kZCta);
    double s0 = 0.0;
#pragma unroll 4
    for (int r = 0; r < G.ncta; ++r) s0 += sl[2 * r];
    ++G.flip;
    return s0;
}
// The pull form of the same sum (same order of additions, same value): the CTA's sum goes into its OWN slot (a local
// store: the release of the cluster barrier has no remote store to drain), after the barrier lanes 0..15 of warp 0 fetch
// the 16 slots with one ld.shared::cluster each and the total is handed to the other warps through shared memory.
__device__ __forceinline__ double z_sum1_pull(ZTeam& G, double a) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = kZT / 32;
    a = warp_sum(a);
    double* sm = reinterpret_cast<double*>(G.dsm) + (G.flip & 1) * 64;
    if (lane == 0) sm[w] = a;
    __syncthreads();
    double* sl = reinterpret_cast<double*>(G.dsm + kOffSlots) + (G.flip & 1) * (2 * kZCta);
    if (threadIdx.x == 0) {
        double ta = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) ta += sm[i];
        sl[0] = ta;
    }
    z_barrier();
    if (w == 0) {
        const double v = (lane < G.ncta) ? z_ld(z_map(z_local(sl), lane)) : 0.0;
        double s0 = 0.0;
        for (int r = 0; r < G.ncta; ++r) s0 += __shfl_sync(0xffffffffu, v, r);
        if (lane == 0) sl[1] = s0;
    }
    __syncthreads();
    const double tot = sl[1];
    ++G.flip;
    return tot;
}
__device__ __forceinline__ double z_sum1(ZTeam& G, double a) {
#ifdef SSN_ZSUM_PULL
    return z_sum1_pull(G, a);
#else
    z_sum1_post(G, a);
    return z_sum1_read(G);
#endif
}

// For the local rows [l0, l1) of level L (2^lt lanes per row): s = sum over the row's entries of M of value * v[column],
// v the vector of the COLUMN level that starts `vb` bytes into every CTA's shared memory; filt = 1 / 2 keeps only the
// entries whose column lies in the second / first segment (the two halves of the bigraph level); gather = false: s = 0.
// rs: the staged row bounds of M (A of this level) or null (row pointers from global memory).
// Then epi(l, row, s) on the row's first lane.  Ends WITHOUT a barrier.
template <class Epi>
__device__ __forceinline__ void z_rows(const ZTeam& G, const ZLevel& L, const ZMat& M, const int* rs, int lt, int l0, int l1, int vb,
                                       int filt, int segb, bool gather, Epi&& epi) {
    const int tpr = 1 << lt, sub = threadIdx.x & (tpr - 1), rpt = kZT >> lt;
    for (int lb = l0; lb < l1; lb += rpt) {
        const int l = lb + (threadIdx.x >> lt);
        const int row = (l < l1) ? z_row(L, G.rank, l) : -1;
        double s = 0.0;
        if (row >= 0 && gather) {
            int e, e1;
            if (rs != nullptr) z_bounds(rs, L, l, e, e1); else { e = M.rp[row]; e1 = M.rp[row + 1]; }
            e += sub;
            for (; e + 3 * tpr < e1; e += 4 * tpr) {
                uint32_t lc[4]; double v[4], xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { lc[u] = M.loc[e + u * tpr]; v[u] = M.cv[e + u * tpr]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[u] = z_keep(lc[u], filt, segb) ? z_gather(G, vb, lc[u]) : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) s = fma(v[u], xv[u], s);
            }
            for (; e < e1; e += tpr) {
                const uint32_t lc = M.loc[e];
                if (z_keep(lc, filt, segb)) s = fma(M.cv[e], z_gather(G, vb, lc), s);
            }
        }
        for (int o = tpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row >= 0 && sub == 0) epi(l, row, s);
    }
}

// The first K entries (per lane) of ONE local row of A, kept in registers across the sweeps of a smoothing loop: value and
// RESOLVED shared-window address of the column's element in vector slot 0 of the level, so a gather is an add and a load.
template <int K>
struct ZRow { zaddr ad[K]; double cv[K]; unsigned mask; int e_more, e_end; };

template <int K>
__device__ __forceinline__ void z_row_load(const ZTeam& G, const ZLevel& L, int lt, int l, bool in_range, int filt, int segb, ZRow<K>& R) {
    const int tpr = 1 << lt, sub = threadIdx.x & (tpr - 1);
    int e0 = 0, e1 = 0;
    if (in_range) z_bounds(z_rs(G, L), L, l, e0, e1);
    e0 += sub;
    R.mask = 0u;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int e = e0 + k * tpr;
        const bool has = e < e1;
        const uint32_t lc = has ? L.A.loc[e] : 0u;
        R.cv[k] = has ? L.A.cv[e] : 0.0;
        R.ad[k] = z_map(G.base + (zaddr)(L.so[0] + (int)(lc & kZOffMask)), (int)(lc >> 24));
        if (has && z_keep(lc, filt, segb)) R.mask |= 1u << k;
    }
    R.e_more = e0 + K * tpr; R.e_end = e1;
}

// the gather table is written once at the start of the launch (before the first cluster barrier) and only read afterwards
#ifdef SSN_ZLOC_NC
#define Z_LDRO(p) __ldg(p)
#else
#define Z_LDRO(p) (*(p))
#endif
// lane-reduced A(row,:)*v from the registers (+ the entries past the K-th from global memory); vrel: byte offset of the
// gathered vector from vector slot 0 of the level.  All K gathers are issued before the first product.
template <int K>
__device__ __forceinline__ double z_row_dot(const ZTeam& G, const ZLevel& L, const ZRow<K>& R, int lt, int vrel, int filt, int segb) {
    const int tpr = 1 << lt;
    double xv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xv[k] = ((R.mask >> k) & 1u) ? z_ld(R.ad[k] + (zaddr)vrel) : 0.0;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) s = fma(R.cv[k], xv[k], s);
    // the entries past the K-th: four at a time, every load of a batch in flight before the first use (one entry per trip
    // was one dependent L2 round trip per entry -- the longest row of the slice set the time of the whole sweep: level 1 of
    // the benchmarked state has 7.8 entries per row on average, 21 at most, and 8 in registers)
    int e = R.e_more;
#pragma unroll 1
    for (; e + 3 * tpr < R.e_end; e += 4 * tpr) {
        uint32_t lc[4]; double v[4], xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { lc[u] = Z_LDRO(L.A.loc + e + u * tpr); v[u] = __ldg(L.A.cv + e + u * tpr); }
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = z_keep(lc[u], filt, segb) ? z_gather(G, L.so[0] + vrel, lc[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) s = fma(v[u], xv[u], s);
    }
    if (e < R.e_end) {
        uint32_t lc[3]; double v[3], xv[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) { const bool has = e + u * tpr < R.e_end; lc[u] = has ? Z_LDRO(L.A.loc + e + u * tpr) : 0u; v[u] = has ? __ldg(L.A.cv + e + u * tpr) : 0.0; }
#pragma unroll
        for (int u = 0; u < 3; ++u) xv[u] = (e + u * tpr < R.e_end && z_keep(lc[u], filt, segb)) ? z_gather(G, L.so[0] + vrel, lc[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 3; ++u) s = fma(v[u], xv[u], s);
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// loc tables of the rows this CTA owns (one thread per row); returns true when an off-diagonal entry joins two rows of
// the same segment (the two-half-sweep smoother needs V and T diagonal)
__device__ __noinline__ bool z_build_loc(const ZTeam& G, const ZLevel& Lrow, const ZLevel& Lcol, const ZMat& M, bool check) {
    bool bad = false;
    if (M.rp == nullptr) return false;
    const int nl = Lrow.rpf + Lrow.rpc;
    for (int l = threadIdx.x; l < nl; l += kZT) {
        const int row = z_row(Lrow, G.rank, l);
        if (row < 0) continue;
        const int e1 = M.rp[row + 1];
        for (int e = M.rp[row]; e < e1; ++e) {
            const int j = M.ci[e];
            M.loc[e] = z_loc(Lcol, j);
            if (check && j != row && ((j < Lcol.Nf) == (row < Lcol.Nf))) bad = true;
        }
    }
    return bad;
}

// ---- halo of the A-gathers.  ld.shared::cluster to the CTA's OWN shared memory runs at the speed of ld.shared (~6 eight-byte
// gathers per cycle), to another CTA at ~0.25 per cycle when the lanes of a warp instruction read scattered words -- and at ~2
// per cycle when they read consecutive words (tools/barrier_bench.py, which = 19 / 20 / 22).  The graphs of the path are local in
// the natural order: a CTA's rows reference few DISTINCT elements of other CTAs' slices.  So each gathered vector slot of a
// level carries, behind the CTA's own slice, a copy of those elements (the halo, ascending column order = runs of consecutive
// remote words): the gather tables of A point into it, every gather of a sweep stays inside the SM, and before a pass that
// gathers slot s the CTA refreshes the halo of s with coalesced remote loads (z_halo_pull; the vector was published by the
// cluster barrier that ended the pass that wrote it, as the remote gathers needed it).  Built once per launch and per CTA: a
// bitmap of the referenced remote columns in scratch shared memory, its popcount prefix = the halo slot of a column.  A CTA
// whose halo does not fit the capacity of the level keeps the owners' locations in its table (hn = 0) -- the decision is
// local, the pulls involve no cluster-wide step.
__device__ __noinline__ void z_build_halo(const ZTeam& G, ZLevel& L, unsigned char* scratch) {
    L.hn = 0;
    const int N = L.N, nW = (N + 31) >> 5;
    if (L.hcap <= 0 || N > 65536 || L.A.rp == nullptr) return;      // bitmap (8 KB) + 16-bit word prefixes (4 KB) <= the 16 KB scratch
    uint32_t* bm = reinterpret_cast<uint32_t*>(scratch);
    unsigned short* pre = reinterpret_cast<unsigned short*>(scratch + 8192);
    int* s_tot = reinterpret_cast<int*>(scratch + 12288);              // [kZT / 32 + 1] (no static shared memory in a cluster kernel:
                                                                       // the host emulation runs the 16 CTAs at the same time)
    for (int w = threadIdx.x; w < nW; w += kZT) bm[w] = 0u;
    __syncthreads();
    const int nl = L.rpf + L.rpc;
    for (int l = threadIdx.x; l < nl; l += kZT) {
        const int row = z_row(L, G.rank, l);
        if (row < 0) continue;
        const int e1 = L.A.rp[row + 1];
        for (int e = L.A.rp[row]; e < e1; ++e) {
            const int j = L.A.ci[e];
            if ((int)(z_loc(L, j) >> 24) != G.rank) atomicOr(&bm[j >> 5], 1u << (j & 31));
        }
    }
    __syncthreads();
    // exclusive prefix of the word popcounts: a contiguous run of words per thread, then a scan of the thread totals
    const int per = (nW + kZT - 1) / kZT, w0 = min((int)threadIdx.x * per, nW), w1 = min(w0 + per, nW);
    int mine = 0;
    for (int w = w0; w < w1; ++w) mine += __popc(bm[w]);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_tot[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) { int acc = 0; for (int i = 0; i < kZT / 32; ++i) { const int t = s_tot[i]; s_tot[i] = acc; acc += t; } s_tot[kZT / 32] = acc; }
    __syncthreads();
    const int H = s_tot[kZT / 32];
    if (H > L.hcap || H == 0) { __syncthreads(); return; }             // uniform over the CTA
    int run = s_tot[wid] + incl - mine;
    uint32_t* hsrc = reinterpret_cast<uint32_t*>(G.dsm + L.hoff);
    for (int w = w0; w < w1; ++w) {
        pre[w] = (unsigned short)run;
        uint32_t bits = bm[w];
        while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1u; hsrc[run++] = z_loc(L, (w << 5) + b); }
    }
    __syncthreads();
    // the table entries of the remote columns now point into the CTA's own halo area
    for (int l = threadIdx.x; l < nl; l += kZT) {
        const int row = z_row(L, G.rank, l);
        if (row < 0) continue;
        const int e1 = L.A.rp[row + 1];
        for (int e = L.A.rp[row]; e < e1; ++e) {
            const int j = L.A.ci[e];
            const uint32_t lc = z_loc(L, j);
            if ((int)(lc >> 24) != G.rank) {
                const int h = (int)pre[j >> 5] + __popc(bm[j >> 5] & ((1u << (j & 31)) - 1u));
                L.A.loc[e] = ((uint32_t)G.rank << 24) | (lc & kZSeg) | (uint32_t)(L.hb + 8 * h);
            }
        }
    }
    __syncthreads();
    L.hn = H;
}

// refresh the halo of vector slot `slot` from the owners' slices (consecutive threads read consecutive halo entries: runs of
// consecutive remote words); the slot must have been published by a cluster barrier.  Ends with a CTA barrier.
__device__ __forceinline__ void z_halo_pull(const ZTeam& G, const ZLevel& L, int slot) {
    if (L.hn == 0) return;
    const uint32_t* hsrc = reinterpret_cast<const uint32_t*>(G.dsm + L.hoff);
    double* halo = reinterpret_cast<double*>(G.dsm + L.so[slot] + L.hb);
    for (int h = threadIdx.x; h < L.hn; h += kZT) halo[h] = z_gather(G, L.so[slot], hsrc[h]);
    __syncthreads();
}

// The smoothing loops for slices that do not fit the register slots (larger systems; the host emulation's small CTAs):
// every sweep re-reads its rows.  Out of line: the cycle's hot code stays small enough for the instruction cache.
struct ZStream { double dot; int cur, flip; };

__device__ __noinline__ ZStream z_jacobi_stream(ZTeam G, const ZLevel& L, int smoth, double rxx, bool ez, double sr, double dotAe, int cur) {
    const int nl = L.rpf + L.rpc;
    const double* r = z_vec(G, L, ZV_R);
    const double* dinv = z_dinv(G, L); const double* Axi = z_axi(G, L);
    for (int s = 0; s < smoth; ++s) {
        const double coef = (sr - dotAe) * rxx;
        const double* ec = z_vec(G, L, cur ? ZV_ALT : ZV_E);
        double* ea = z_vec(G, L, cur ? ZV_E : ZV_ALT);
        double part = 0.0;
        if (!ez) z_halo_pull(G, L, cur ? ZV_ALT : ZV_E);
        z_rows(G, L, L.A, z_rs(G, L), L.ltA, 0, nl, z_vb(L, cur ? ZV_ALT : ZV_E), 0, 0, !ez, [&](int l, int, double d) {
            const double axi = Axi[l], ei = ez ? 0.0 : ec[l];
            const double en = ei + coef + dinv[l] * ((r[l] - d) - axi * coef);
            ea[l] = en;
            part = fma(axi, en, part);
        });
        dotAe = z_sum1(G, part); cur ^= 1; ez = false;
    }
    return ZStream{dotAe, cur, G.flip};
}

__device__ __noinline__ ZStream z_gs_stream(ZTeam G, const ZLevel& L, int smoth, double rxx, bool post, bool ez, double sr, double dotAe) {
    const double* r = z_vec(G, L, ZV_R);
    const double* dinv = z_dinv(G, L); const double* Axi = z_axi(G, L);
    double* e = z_vec(G, L, ZV_E);
    const int vbe = z_vb(L, ZV_E), segb = L.rpf * 8;
    const int a0 = post ? L.rpf : 0, a1 = post ? L.rpf + L.rpc : L.rpf;
    const int b0 = post ? 0 : L.rpf, b1 = post ? L.rpf : L.rpf + L.rpc;
    const int fa = post ? 2 : 1, fb = post ? 1 : 2;
    for (int s = 0; s < smoth; ++s) {
        const double coef = (sr - dotAe) * rxx;
        double part = 0.0;
        if (!ez) z_halo_pull(G, L, ZV_E);
        z_rows(G, L, L.A, z_rs(G, L), L.ltG, a0, a1, vbe, fa, segb, !ez, [&](int l, int, double d) {
            const double axi = Axi[l];
            const double en = coef + dinv[l] * ((r[l] - d) - axi * coef);
            e[l] = en;
            part = fma(axi, en, part);
        });
        z_barrier();
        z_halo_pull(G, L, ZV_E);
        z_rows(G, L, L.A, z_rs(G, L), L.ltG, b0, b1, vbe, fb, segb, true, [&](int l, int, double d) {
            const double en = dinv[l] * (r[l] - d);
            e[l] = en;
            part = fma(Axi[l], en, part);
        });
        dotAe = z_sum1(G, part); ez = false;
    }
    return ZStream{dotAe, 0, G.flip};
}

// (a function of its own, like the streaming smoothers: the solve kernel's register allocation stays what it was)
// e = PCG(A_k, r) on the leaf level of the two-level method: zero guess, Jacobi preconditioner, stops on
// delta_new <= tol^2 * delta_0 or after pcg_maxit iterations                                         PCG.m:68-88, twogrid_bigph.m:98-99
// (the arithmetic of pcg_kernel, amg_solve.cu; p in ALT, q / w in G, the residual in slot 4, the solution in E).  Per iteration:
// one gathering pass over p, two cluster-wide sums and one plain barrier (p complete before it is gathered again).
__device__ __noinline__ int z_pcg_leaf(ZTeam G, const ZLevel& L, int pcg_maxit, double pcg_tol2, bool noreg, int tail_off, int tail_cap) {
    const int nl = L.rpf + L.rpc;
    double* d = z_vec(G, L, ZV_E);
    const double* rhs = z_vec(G, L, ZV_R);
    double* q = z_vec(G, L, ZV_G);
    double* p = z_vec(G, L, ZV_ALT);
    double* rr = z_vec(G, L, 4);
    const double* dg = z_dinv(G, L);                                // diag(A_k) here (leaf_pcg)
    double dn = 0.0;
    for (int l = threadIdx.x; l < nl; l += kZT) {
        if (z_row(L, G.rank, l) < 0) continue;
        const double ri = rhs[l], wi = ri / dg[l];
        d[l] = 0.0; rr[l] = ri; p[l] = wi; dn = fma(ri, wi, dn);
    }
    double delta_new = z_sum1(G, dn);
    const double delta_0 = delta_new;
    const bool regs = !noreg && (nl << L.ltA) <= kZSlots * kZT;
    constexpr int K = 8;
    const bool first = (threadIdx.x & ((1 << L.ltA) - 1)) == 0;
    ZRow<K> R[kZSlots];
    int lr[kZSlots]; bool mine[kZSlots];
    if (regs) {
#pragma unroll
        for (int u = 0; u < kZSlots; ++u) {
            lr[u] = (threadIdx.x >> L.ltA) + u * (kZT >> L.ltA);
            mine[u] = lr[u] < nl && z_row(L, G.rank, lr[u]) >= 0;
            z_row_load<K>(G, L, L.ltA, lr[u], mine[u], 0, 0, R[u]);
        }
    }
    // The entries of a row past the K in registers: one lane per row (the usual case at these row lengths) copies them into
    // shared memory once per solve -- column location and value, in order -- so that the ~100 passes of the loop below do not
    // fetch them from L2 (two dependent round trips per batch of a long row: the slowest thread sets the time of a pass).
    // Same products in the same order as z_row_dot.  Rows that do not fit the area keep reading global memory.
    int tb[kZSlots], tn[kZSlots];
#pragma unroll
    for (int u = 0; u < kZSlots; ++u) { tb[u] = -1; tn[u] = 0; }
    double* tv = reinterpret_cast<double*>(G.dsm + tail_off);
    uint32_t* tl = reinterpret_cast<uint32_t*>(G.dsm + tail_off + (size_t)tail_cap * 8);
    if (regs && L.ltA == 0 && tail_cap > 0) {
        int* tc = reinterpret_cast<int*>(G.dsm + kOffTailCnt);
        if (threadIdx.x == 0) *tc = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kZSlots; ++u) {
            const int len = mine[u] ? R[u].e_end - R[u].e_more : 0;
            if (len > 0) {
                const int base = atomicAdd(tc, len);
                if (base + len <= tail_cap) {
                    for (int j = 0; j < len; ++j) { tl[base + j] = L.A.loc[R[u].e_more + j]; tv[base + j] = L.A.cv[R[u].e_more + j]; }
                    tb[u] = base; tn[u] = len; R[u].e_end = R[u].e_more;
                }
            }
        }
        __syncthreads();
    }
    // One lane per row and every row of the slice in a register slot: the thread that gathers for a row is the thread that
    // updates it, so the row's d, r, own p and diagonal stay in registers for the whole solve (q and w are never stored; only p
    // goes to shared memory, for the other rows' gathers).  Same expressions, same order of the partial sums as the loops below.
    const bool own = regs && L.ltA == 0;
    double dR[kZSlots], rR[kZSlots], pR[kZSlots], gR[kZSlots];
#pragma unroll
    for (int u = 0; u < kZSlots; ++u) {
        const bool on = own && mine[u];
        dR[u] = 0.0; rR[u] = on ? rr[lr[u]] : 0.0; pR[u] = on ? p[lr[u]] : 0.0; gR[u] = on ? dg[lr[u]] : 1.0;
    }
    int it = 0;
    // debug build: cycles of the five phases of an iteration (lead thread), slots 8..12 of the 'zsum/dots' row of g_zdbg
#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
#define ZPH(slot) do { if (G.rank == 0 && threadIdx.x == 0) { const long long t__ = clock64(); g_zdbg[6 * 16 + (slot)] += (unsigned long long)(t__ - tph); g_zdbg[128 + 6 * 16 + (slot)] += 1ull; tph = t__; } } while (0)
    long long tph = clock64();
#else
#define ZPH(slot) do {} while (0)
#endif
    while (it < pcg_maxit && delta_new > pcg_tol2 * delta_0) {  // PCG.m:76
        const double delta_old = delta_new;
        double qp = 0.0;
        if (regs) {
            double sv[kZSlots];
#pragma unroll
            for (int u = 0; u < kZSlots; ++u) sv[u] = z_row_dot<K>(G, L, R[u], L.ltA, L.so[ZV_ALT] - L.so[0], 0, 0);
#pragma unroll
            for (int u = 0; u < kZSlots; ++u) {
                int e = 0;
                for (; e + 3 < tn[u]; e += 4) {
                    double xv[4];
#pragma unroll
                    for (int w = 0; w < 4; ++w) xv[w] = z_gather(G, L.so[ZV_ALT], tl[tb[u] + e + w]);
#pragma unroll
                    for (int w = 0; w < 4; ++w) sv[u] = fma(tv[tb[u] + e + w], xv[w], sv[u]);
                }
                for (; e < tn[u]; ++e) sv[u] = fma(tv[tb[u] + e], z_gather(G, L.so[ZV_ALT], tl[tb[u] + e]), sv[u]);
            }
            if (own) {
#pragma unroll
                for (int u = 0; u < kZSlots; ++u) if (mine[u]) qp = fma(sv[u], pR[u], qp);
                ZPH(8);
                qp = z_sum1(G, qp);
                ZPH(9);
                const double alpha = delta_old / qp;
                double wR[kZSlots];
                dn = 0.0;
#pragma unroll
                for (int u = 0; u < kZSlots; ++u) {
                    wR[u] = 0.0;
                    if (mine[u]) {
                        dR[u] += alpha * pR[u];
                        const double ri = rR[u] - alpha * sv[u];
                        const double wi = ri / gR[u];
                        rR[u] = ri; wR[u] = wi; dn = fma(ri, wi, dn);
                    }
                }
                ZPH(10);
                delta_new = z_sum1(G, dn);
                ZPH(11);
                const double beta = delta_new / delta_old;
#pragma unroll
                for (int u = 0; u < kZSlots; ++u) if (mine[u]) { pR[u] = wR[u] + beta * pR[u]; p[lr[u]] = pR[u]; }
                ++it;
                z_barrier();
                ZPH(12);
                continue;
            }
#pragma unroll
            for (int u = 0; u < kZSlots; ++u) if (mine[u] && first) { q[lr[u]] = sv[u]; qp = fma(sv[u], p[lr[u]], qp); }
        } else {
            z_rows(G, L, L.A, z_rs(G, L), L.ltA, 0, nl, z_vb(L, ZV_ALT), 0, 0, true, [&](int l, int, double sdot) { q[l] = sdot; qp = fma(sdot, p[l], qp); });
        }
        ZPH(8);
        qp = z_sum1(G, qp);                                         // every gather of p is done behind its barrier
        ZPH(9);
        const double alpha = delta_old / qp;
        dn = 0.0;
        for (int l = threadIdx.x; l < nl; l += kZT) {
            if (z_row(L, G.rank, l) < 0) continue;
            d[l] += alpha * p[l];
            const double ri = rr[l] - alpha * q[l];
            const double wi = ri / dg[l];
            rr[l] = ri; q[l] = wi; dn = fma(ri, wi, dn);
        }
        ZPH(10);
        delta_new = z_sum1(G, dn);
        ZPH(11);
        const double beta = delta_new / delta_old;
        for (int l = threadIdx.x; l < nl; l += kZT) if (z_row(L, G.rank, l) >= 0) p[l] = q[l] + beta * p[l];
        ++it;
        z_barrier();
        ZPH(12);
    }
#undef ZPH
    if (own) {
#pragma unroll
        for (int u = 0; u < kZSlots; ++u) if (mine[u]) d[lr[u]] = dR[u];
        z_barrier();                                        // the prolongation gathers the solution from every CTA
    }
    return G.flip;
}

#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
#define ZDBG(op, level, call) do { const long long t0__ = clock64(); call; if (lead) { g_zdbg[(op) * 16 + (level)] += (unsigned long long)(clock64() - t0__); g_zdbg[128 + (op) * 16 + (level)] += 1ull; } } while (0)
#else
#define ZDBG(op, level, call) do { call; } while (0)
#endif

}  // namespace

void debug_cycles_dsm(unsigned long long* out256, bool reset) {
#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
    cudaMemcpyFromSymbol(out256, g_zdbg, sizeof(unsigned long long) * 256);
    if (reset) { unsigned long long z[256] = {0}; cudaMemcpyToSymbol(g_zdbg, z, sizeof(z)); }
#else
    for (int i = 0; i < 256; ++i) out256[i] = 0ull;
    (void)reset;
#endif
}

#ifdef SSN_EMU
void dsm_solve_kernel(const ZArgs a) {
    unsigned char* dsm = emu::cluster_smem[blockIdx.x];
#else
__global__ void __launch_bounds__(kZT, 1) dsm_solve_kernel(const __grid_constant__ ZArgs a) {
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    ZTeam G{z_rank(), z_ncta(), 0, dsm, z_local(dsm)};
    const bool lead = (G.rank == 0 && threadIdx.x == 0);
    ZLevel* sl = reinterpret_cast<ZLevel*>(dsm + kOffLv);
    int* prog = reinterpret_cast<int*>(dsm + kOffProg);
    double* st_sum_r = reinterpret_cast<double*>(dsm + kOffSumR);
    double* st_dot = reinterpret_cast<double*>(dsm + kOffDot);
    int* st_cur = reinterpret_cast<int*>(dsm + kOffCur);
    double* xs = reinterpret_cast<double*>(dsm + kOffXs);
    const int kd = a.kd;
    for (int t = threadIdx.x; t <= kd; t += kZT) sl[t] = a.lv[t];
    for (int t = threadIdx.x; t < a.nprog; t += kZT) prog[t] = a.prog[t];
    for (int t = threadIdx.x; t < kZMaxL; t += kZT) { st_sum_r[t] = 0.0; st_dot[t] = 0.0; st_cur[t] = 0; }
    __syncthreads();
    // ---- gather tables; the structural requirement of the bigraph smoother
    bool bad = false;
    for (int k = 0; k <= kd; ++k) {
        bad |= z_build_loc(G, sl[k], sl[k], sl[k].A, k == 0 && sl[0].bigph != 0);
        if (k < kd) z_build_loc(G, sl[k], sl[k + 1], sl[k].Pu, false);
        if (k >= 1) z_build_loc(G, sl[k], sl[k - 1], sl[k].Td, false);
    }
    // ---- halos of the A-gathers (each CTA's own decision per level; the scratch is the dense leaf's vector buffer)
    __syncthreads();                                                    // L.hn is written by all threads with the same value: sl is shared
    for (int k = 0; k <= kd; ++k) z_build_halo(G, sl[k], dsm + kOffXs);
    // ---- row bounds of A, 1/diag, A*ones of this CTA's rows: shared memory
    for (int k = 0; k <= kd; ++k) {
        const ZLevel& L = sl[k];
        int* rs = reinterpret_cast<int*>(dsm + L.poff);
        const int f0 = min(G.rank * L.rpf, L.Nf), f1 = min(L.Nf, f0 + L.rpf);
        const int c0 = min(L.Nf + G.rank * L.rpc, L.N), c1 = min(L.N, c0 + L.rpc);
        for (int l = threadIdx.x; l <= L.rpf; l += kZT) rs[l] = L.A.rp[min(f0 + l, f1)];
        for (int l = threadIdx.x; l <= L.rpc; l += kZT) rs[L.rpf + 1 + l] = L.A.rp[min(c0 + l, c1)];
        if (k < kd) {
            double* di = reinterpret_cast<double*>(dsm + L.poff + L.rsbytes);
            double* ax = di + (L.rpf + L.rpc);
            for (int l = threadIdx.x; l < L.rpf + L.rpc; l += kZT) {
                const int row = z_row(L, G.rank, l);
                di[l] = row >= 0 ? L.dinv[row] : 0.0; ax[l] = row >= 0 ? L.Axi[row] : 0.0;
            }
        } else if (a.leaf_pcg) {
            // the PCG leaf divides by the diagonal itself (PCG.m:93, as pcg_kernel does): diag(A_kd) of this CTA's rows
            double* dg = reinterpret_cast<double*>(dsm + L.poff + L.rsbytes);
            for (int l = threadIdx.x; l < L.rpf + L.rpc; l += kZT) {
                const int row = z_row(L, G.rank, l);
                double dv = 0.0;
                if (row >= 0) for (int e = L.A.rp[row]; e < L.A.rp[row + 1]; ++e) if (L.A.ci[e] == row) dv = L.A.cv[e];
                dg[l] = dv;
            }
        }
    }
    const ZLevel& L0 = sl[0];
    const int nl0 = L0.rpf + L0.rpc;
    double* x0 = z_vec(G, L0, L0.xslot);
    double* r0 = z_vec(G, L0, ZV_R);
    for (int l = threadIdx.x; l < nl0; l += kZT) { const int row = z_row(L0, G.rank, l); if (row >= 0) x0[l] = a.x[row]; }
    {
        double fb = bad ? 1.0 : 0.0, dummy = 0.0;
        z_sum2(G, fb, dummy);                                           // also: x slices and loc tables are in place
        if (fb != 0.0) {
            if (G.rank == 0 && threadIdx.x == 0) { a.it_out[0] = 0; a.it_out[1] = 1; a.it_out[2] = 1; }
            z_barrier();
            return;
        }
    }
    const int smoth = a.smoth;
    const bool nsp = a.isnsp != 0;

    // r = b - A*x ; sum(r), sum(r^2)                                    Class_AMG.m:89 / :96,:102
    auto outer_residual = [&](double& s1, double& s2) {
        s1 = 0.0; s2 = 0.0;
        z_halo_pull(G, L0, L0.xslot);
        z_rows(G, L0, L0.A, z_rs(G, L0), L0.ltA, 0, nl0, z_vb(L0, L0.xslot), 0, 0, true, [&](int l, int row, double s) {
            const double ri = a.b[row] - s; r0[l] = ri; s1 += ri; s2 = fma(ri, ri, s2);
        });
        z_sum2(G, s1, s2);
    };
    // smoth damped-Jacobi sweeps on level k (ping-pong E <-> ALT); returns Axi'e of the result        MG_Wcycle.m:15-23
    auto jacobi = [&](int k, bool ez, double sr, double dotAe) -> double {
        const ZLevel& L = sl[k];
        const int nl = L.rpf + L.rpc;
        const double* r = z_vec(G, L, ZV_R);
        const double* dinv = z_dinv(G, L); const double* Axi = z_axi(G, L);
        // the kernel-component coefficient xi'g / xx as a product with 1/xx (one rounding of difference, a division off the
        // critical path of every sweep); 0 when the system is not treated as nearly singular
        const double rxx = nsp ? 1.0 / L.xx : 0.0;
        int cur = st_cur[k];
        if (!a.noreg && (nl << L.ltA) <= kZSlots * kZT) {
            // every row of the slice has its own lanes: the row's entries stay in registers for all the sweeps
            constexpr int K = 8;
            const bool first = (threadIdx.x & ((1 << L.ltA) - 1)) == 0;
            ZRow<K> R[kZSlots];
            int l[kZSlots]; bool mine[kZSlots]; double axi[kZSlots], di[kZSlots], ri[kZSlots];
#pragma unroll
            for (int u = 0; u < kZSlots; ++u) {
                l[u] = (threadIdx.x >> L.ltA) + u * (kZT >> L.ltA);
                mine[u] = l[u] < nl && z_row(L, G.rank, l[u]) >= 0;
                z_row_load<K>(G, L, L.ltA, l[u], mine[u], 0, 0, R[u]);
                axi[u] = mine[u] ? Axi[l[u]] : 0.0; di[u] = mine[u] ? dinv[l[u]] : 0.0; ri[u] = mine[u] ? r[l[u]] : 0.0;
            }
            for (int s = 0; s < smoth; ++s) {
                const double coef = (sr - dotAe) * rxx;
                const double* ec = z_vec(G, L, cur ? ZV_ALT : ZV_E);
                double* ea = z_vec(G, L, cur ? ZV_E : ZV_ALT);
                double d[kZSlots];
                if (!ez) z_halo_pull(G, L, cur ? ZV_ALT : ZV_E);
#pragma unroll
                for (int u = 0; u < kZSlots; ++u) d[u] = ez ? 0.0 : z_row_dot<K>(G, L, R[u], L.ltA, L.so[cur ? ZV_ALT : ZV_E] - L.so[0], 0, 0);
                double part = 0.0;
#pragma unroll
                for (int u = 0; u < kZSlots; ++u)
                    if (mine[u] && first) {
                        const double ei = ez ? 0.0 : ec[l[u]];
                        const double en = ei + coef + di[u] * ((ri[u] - d[u]) - axi[u] * coef);
                        ea[l[u]] = en;
                        part = fma(axi[u], en, part);
                    }
                dotAe = z_sum1(G, part); cur ^= 1; ez = false;
            }
        } else {
            const ZStream o = z_jacobi_stream(G, L, smoth, rxx, ez, sr, dotAe, cur);
            dotAe = o.dot; cur = o.cur; G.flip = o.flip;
        }
        st_cur[k] = cur;
        return dotAe;
    };
    // smoth block Gauss-Seidel sweeps on the bigraph level 0, in place in E; returns Axi'e              Class_AMG.m:48-59
    auto gauss_seidel = [&](bool post, bool ez, double sr, double dotAe) -> double {
        const ZLevel& L = L0;
        const double* r = r0;
        const double* dinv = z_dinv(G, L); const double* Axi = z_axi(G, L);
        const double rxx = nsp ? 1.0 / L.xx : 0.0;
        double* e = z_vec(G, L, ZV_E);
        const int vbe = z_vb(L, ZV_E), segb = L.rpf * 8;
        // pre: first-segment rows with the kernel correction, then second-segment rows; post: the other way round
        const int a0 = post ? L.rpf : 0, a1 = post ? L.rpf + L.rpc : L.rpf;
        const int b0 = post ? 0 : L.rpf, b1 = post ? L.rpf : L.rpf + L.rpc;
        const int fa = post ? 2 : 1, fb = post ? 1 : 2;
        if (!a.noreg && (max(L.rpf, L.rpc) << L.ltG) <= kZSlots * kZT) {
            constexpr int K = 4;
            const bool first = (threadIdx.x & ((1 << L.ltG) - 1)) == 0;
            ZRow<K> Ra[kZSlots], Rb[kZSlots];
            int la[kZSlots], lb[kZSlots]; bool ma[kZSlots], mb[kZSlots];
            double axa[kZSlots], dia[kZSlots], ra[kZSlots], axb[kZSlots], dib[kZSlots], rb[kZSlots];
#pragma unroll
            for (int u = 0; u < kZSlots; ++u) {
                const int lq = (threadIdx.x >> L.ltG) + u * (kZT >> L.ltG);
                la[u] = a0 + lq; lb[u] = b0 + lq;
                ma[u] = la[u] < a1 && z_row(L, G.rank, la[u]) >= 0; mb[u] = lb[u] < b1 && z_row(L, G.rank, lb[u]) >= 0;
                z_row_load<K>(G, L, L.ltG, la[u], ma[u], fa, segb, Ra[u]);
                z_row_load<K>(G, L, L.ltG, lb[u], mb[u], fb, segb, Rb[u]);
                axa[u] = ma[u] ? Axi[la[u]] : 0.0; dia[u] = ma[u] ? dinv[la[u]] : 0.0; ra[u] = ma[u] ? r[la[u]] : 0.0;
                axb[u] = mb[u] ? Axi[lb[u]] : 0.0; dib[u] = mb[u] ? dinv[lb[u]] : 0.0; rb[u] = mb[u] ? r[lb[u]] : 0.0;
            }
            for (int s = 0; s < smoth; ++s) {
                const double coef = (sr - dotAe) * rxx;
                double part = 0.0;
                double d[kZSlots];
                if (!ez) z_halo_pull(G, L, ZV_E);
#pragma unroll
                for (int u = 0; u < kZSlots; ++u) d[u] = ez ? 0.0 : z_row_dot<K>(G, L, Ra[u], L.ltG, 0, fa, segb);
#pragma unroll
                for (int u = 0; u < kZSlots; ++u)
                    if (ma[u] && first) { const double en = coef + dia[u] * ((ra[u] - d[u]) - axa[u] * coef); e[la[u]] = en; part = fma(axa[u], en, part); }
                z_barrier();
                z_halo_pull(G, L, ZV_E);
#pragma unroll
                for (int u = 0; u < kZSlots; ++u) d[u] = z_row_dot<K>(G, L, Rb[u], L.ltG, 0, fb, segb);
#pragma unroll
                for (int u = 0; u < kZSlots; ++u)
                    if (mb[u] && first) { const double en = dib[u] * (rb[u] - d[u]); e[lb[u]] = en; part = fma(axb[u], en, part); }
                dotAe = z_sum1(G, part); ez = false;
            }
        } else {
            const ZStream o = z_gs_stream(G, L, smoth, rxx, post, ez, sr, dotAe);
            dotAe = o.dot; G.flip = o.flip;
        }
        return dotAe;
    };

#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
    const long long t_kernel = clock64();
#endif
    // the state of the outer loop lives in shared memory while a cycle runs (every thread holds the same values and writes
    // them; registers are for the rows of the smoothing loops): [0] res0, [1] res_prev, [2] rel_prev
    double* st_outer = reinterpret_cast<double*>(dsm + kOffOuter);
    int it = 0, hist = 1;
    {
        double s1, s2;
        outer_residual(s1, s2);
        const double res0 = sqrt(s2);
        st_outer[0] = res0; st_outer[1] = res0; st_outer[2] = 1.0; st_sum_r[0] = s1;
        if (lead) { a.relk[0] = 1.0; a.rho[0] = NAN; a.it_out[2] = 0; }
        if (res0 == 0.0) {
            if (lead) { a.relk[0] = 0.0; a.rho[0] = INFINITY; a.it_out[0] = 0; a.it_out[1] = 1; }
            z_barrier();
            return;
        }
    }
    it = 1;
    while (st_outer[2] > a.retol && it <= a.maxit) {                    // Class_AMG.m:95
        for (int pc = 0; pc < a.nprog; ++pc) {
            const int op = prog[pc] & 0xff, k = (prog[pc] >> 8) & 0xff;
            const bool zero = ((prog[pc] >> 16) & 1) != 0;
            const ZLevel& L = sl[k];
            const int nl = L.rpf + L.rpc;
            if (op == Z_PRE || op == Z_POST) {
                const bool post = (op == Z_POST);
                const double sr = st_sum_r[k];
                const double d0 = (op == Z_PRE && zero) ? 0.0 : st_dot[k];
                if (op == Z_PRE && zero) st_cur[k] = 0;
                double d;
                if (k == 0 && L.bigph) ZDBG(1, k, d = gauss_seidel(post, op == Z_PRE && zero, sr, d0));
                else ZDBG(2, k, d = jacobi(k, op == Z_PRE && zero, sr, d0));
                st_dot[k] = d;
            } else if (op == Z_RESTRICT) {                              // r_{k+1} = Pro' (r - A e)            MG_Wcycle.m:26
                const double* r = z_vec(G, L, ZV_R);
                double* g = z_vec(G, L, ZV_G);
                const int es = (k == 0 && L.bigph) ? ZV_E : (st_cur[k] ? ZV_ALT : ZV_E);
                ZDBG(0, k, z_halo_pull(G, L, es);
                z_rows(G, L, L.A, z_rs(G, L), L.ltA, 0, nl, z_vb(L, es), 0, 0, true, [&](int l, int, double d) { g[l] = r[l] - d; });
                z_barrier());
                const ZLevel& Lc = sl[k + 1];
                double* rc = z_vec(G, Lc, ZV_R);
                double sy = 0.0;
                ZDBG(3, k, z_rows(G, Lc, Lc.Td, nullptr, Lc.ltT, 0, Lc.rpf + Lc.rpc, z_vb(L, ZV_G), 0, 0, true, [&](int l, int, double d) { rc[l] = d; sy += d; });
                sy = z_sum1(G, sy));
                st_sum_r[k + 1] = sy;
            } else if (op == Z_PROLONG) {                               // e += Pro e_{k+1}, with Axi'e        MG_Wcycle.m:32
                const ZLevel& Lc = sl[k + 1];
                const int cs = (k + 1 == kd) ? ZV_E : (st_cur[k + 1] ? ZV_ALT : ZV_E);
                const int es = (k == 0 && L.bigph) ? ZV_E : (st_cur[k] ? ZV_ALT : ZV_E);
                double* e = z_vec(G, L, es);
                double swy = 0.0;
                const double* Axi = z_axi(G, L);
                ZDBG(3, k, z_rows(G, L, L.Pu, nullptr, L.ltP, 0, nl, z_vb(Lc, cs), 0, 0, true, [&](int l, int, double d) {
                    const double v = e[l] + d; e[l] = v; swy = fma(Axi[l], v, swy);
                });
                swy = z_sum1(G, swy));
                st_dot[k] = swy;
            } else if (op == Z_PCG) {                                   // e = PCG(A_k, r)                      twogrid_bigph.m:99
                ZDBG(4, k, G.flip = z_pcg_leaf(G, L, a.pcg_maxit, a.pcg_tol2, a.noreg != 0, a.tail_off, a.tail_cap));
            } else {                                                    // Z_LEAF: e = B r, or e += B (r - A e)
                const double* r = z_vec(G, L, ZV_R);
                double* e = z_vec(G, L, ZV_E);
                int in_slot = ZV_R;
                if (!zero) {
                    double* g = z_vec(G, L, ZV_G);
                    ZDBG(0, k, z_halo_pull(G, L, ZV_E);
                    z_rows(G, L, L.A, z_rs(G, L), L.ltA, 0, nl, z_vb(L, ZV_E), 0, 0, true, [&](int l, int, double d) { g[l] = r[l] - d; });
                    z_barrier());
                    in_slot = ZV_G;
                }
                const int n = L.N, vb = z_vb(L, in_slot);
#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
                const long long t_leaf = clock64();
#endif
                for (int j = threadIdx.x; j < n; j += kZT) xs[j] = z_gather(G, vb, z_loc(L, j));
                __syncthreads();
                const int lane = threadIdx.x & 31;
                // a warp owns up to RW rows of the slice and streams them TOGETHER: RW x 8 loads of B in flight per lane
                // instead of 8 (the rows are read from L2; one row after the other was 9 dependent round trips at N = 651)
                constexpr int RW = 3, NWZ = kZT / 32;
                for (int l0 = (threadIdx.x >> 5); l0 < nl; l0 += RW * NWZ) {
                    int rowv[RW]; const double* Br[RW]; double acc[RW][4];
#pragma unroll
                    for (int w = 0; w < RW; ++w) {
                        const int l = l0 + w * NWZ;
                        rowv[w] = (l < nl) ? z_row(L, G.rank, l) : -1;
                        Br[w] = L.B + (size_t)(rowv[w] >= 0 ? rowv[w] : 0) * n;
                        acc[w][0] = acc[w][1] = acc[w][2] = acc[w][3] = 0.0;
                    }
                    for (int j0 = 0; j0 < n; j0 += 256) {
                        double bv[RW][8], xv[8];
#pragma unroll
                        for (int w = 0; w < RW; ++w)
#pragma unroll
                            for (int u = 0; u < 8; ++u) { const int j = j0 + u * 32 + lane; bv[w][u] = (j < n && rowv[w] >= 0) ? Br[w][j] : 0.0; }
#pragma unroll
                        for (int u = 0; u < 8; ++u) { const int j = j0 + u * 32 + lane; xv[u] = (j < n) ? xs[j] : 0.0; }
#pragma unroll
                        for (int w = 0; w < RW; ++w)
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc[w][u & 3] = fma(bv[w][u], xv[u], acc[w][u & 3]);
                    }
#pragma unroll
                    for (int w = 0; w < RW; ++w) {
                        const double sdot = warp_sum((acc[w][0] + acc[w][1]) + (acc[w][2] + acc[w][3]));
                        const int l = l0 + w * NWZ;
                        if (lane == 0 && rowv[w] >= 0) e[l] = zero ? sdot : (e[l] + sdot);
                    }
                }
                z_barrier();
#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
                if (lead) { g_zdbg[4 * 16 + k] += (unsigned long long)(clock64() - t_leaf); g_zdbg[128 + 4 * 16 + k] += 1ull; }
#endif
            }
        }
        // ---------------- x += e ; r = b - A*x ; res = norm(r)          Class_AMG.m:96-104
        {
            const double* et = z_vec(G, L0, L0.bigph ? ZV_E : (st_cur[0] ? ZV_ALT : ZV_E));
            for (int l = threadIdx.x; l < nl0; l += kZT) x0[l] += et[l];
            z_barrier();
        }
        double s1, s2;
        ZDBG(5, 0, outer_residual(s1, s2));
        st_sum_r[0] = s1;
        const double res = sqrt(s2);
        const double rel_res = res / st_outer[0], rho = res / st_outer[1];
        if (lead) { a.relk[it] = rel_res; a.rho[it] = rho; }
        __syncthreads();                                                // every thread has read res_prev before it changes
        st_outer[1] = res; st_outer[2] = rel_res;
        ++it; ++hist;
        if (rho > 1.0) break;                                           // Class_AMG.m:106
    }
    for (int l = threadIdx.x; l < nl0; l += kZT) { const int row = z_row(L0, G.rank, l); if (row >= 0) a.x[row] = x0[l]; }
    if (lead) { a.it_out[0] = it - 1; a.it_out[1] = hist; int hs = 0; for (int k = 0; k <= kd; ++k) hs += sl[k].hn; a.it_out[3] = hs; }   // [3]: halo entries of CTA 0
#if defined(SSN_PERSIST_DEBUG) && !defined(SSN_EMU)
    if (lead) { g_zdbg[7 * 16] += (unsigned long long)(clock64() - t_kernel); g_zdbg[128 + 7 * 16] += 1ull; }
#endif
    z_barrier();                                            // no CTA exits while a peer may still read its shared memory
}

namespace {

#ifndef SSN_EMU
// ---- micro-benchmarks of the building blocks of a pass (development aid, ssn_debug_barrier_bench which >= 10): one
// cluster of 16 CTAs x kZT threads, `iters` iterations of
//   10 z_sum1   11 z_barrier   12 NG gathers through ld.shared::cluster, all in flight, + barrier
//   13 the same gathers in dependent batches of 4   14 NG gathers from global memory (ld.global.cg) + barrier
//   15 NG gathers from the CTA's own shared memory + barrier   16 z_sum1 without its cluster barrier's remote stores
//      (block reduction + barrier)   17 local store + relaxed-arrive barrier   18 z_sum1_pull
// ng = gathers per thread (<= 16); the gathered vector has 16 * 1024 doubles, indices pseudo-random
__global__ void __launch_bounds__(kZT, 1) dsm_bench_kernel(double* gbuf, int iters, int which, int ng, long long* cycles_out) {
    extern __shared__ __align__(16) unsigned char dsm[];
    ZTeam G{z_rank(), z_ncta(), 0, dsm, z_local(dsm)};
    double* v = reinterpret_cast<double*>(dsm + 4096);
    for (int i = threadIdx.x; i < 1024; i += kZT) v[i] = 1.0 + i;
    uint32_t lc[16]; int gi[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        const unsigned h = (unsigned)(G.rank * kZT + threadIdx.x) * 2654435761u + (unsigned)u * 40503u;
        int idx = (int)((h >> 7) % (16u * 1024u));
        // 19: every gather goes to the CTA's OWN slice through ld.shared::cluster; 20: to the next CTA's; 21: 3 of 4 own, 1 of 4 next
        if (which == 19) idx = (G.rank << 10) | (idx & 1023);
        if (which == 20) idx = (((G.rank + 1) & 15) << 10) | (idx & 1023);
        if (which == 21) idx = ((((u & 3) == 3 ? G.rank + 1 : G.rank) & 15) << 10) | (idx & 1023);
        // 22: coalesced remote loads (consecutive lanes read consecutive doubles of the next CTA's slice); 23: the same from 4 owners
        if (which == 22) idx = (((G.rank + 1) & 15) << 10) | ((threadIdx.x + 512 * u) & 1023);
        if (which == 23) idx = (((G.rank + 1 + (u & 3)) & 15) << 10) | ((threadIdx.x + 512 * (u >> 2)) & 1023);
        gi[u] = idx; lc[u] = ((uint32_t)(idx >> 10) << 24) | (uint32_t)((idx & 1023) * 8);
    }
    z_barrier();
    const long long t0 = clock64();
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        if (which == 10) acc += z_sum1(G, acc + threadIdx.x);
        else if (which == 18) acc += z_sum1_pull(G, acc + threadIdx.x);
        else if (which == 11) z_barrier();
        else if (which == 12 || which == 13 || (which >= 19 && which <= 23)) {
            double xv[16];
            if (which != 13) {
#pragma unroll
                for (int u = 0; u < 16; ++u) xv[u] = (u < ng) ? z_gather(G, 4096, lc[u]) : 0.0;
#pragma unroll
                for (int u = 0; u < 16; ++u) acc = fma(1e-9, xv[u], acc);
            } else {
#pragma unroll
                for (int u0 = 0; u0 < 16; u0 += 4) {
#pragma unroll
                    for (int u = u0; u < u0 + 4; ++u) xv[u] = (u < ng) ? z_gather(G, 4096, lc[u]) : 0.0;
#pragma unroll
                    for (int u = u0; u < u0 + 4; ++u) acc = fma(1e-9, xv[u], acc);
                }
            }
            v[threadIdx.x] = acc;
            z_barrier();
        } else if (which == 14) {
            double xv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) xv[u] = (u < ng) ? __ldcg(gbuf + gi[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) acc = fma(1e-9, xv[u], acc);
            gbuf[G.rank * 1024 + threadIdx.x] = acc;
            z_barrier();
        } else if (which == 15) {
            double xv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) xv[u] = (u < ng) ? v[gi[u] & 1023] : 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) acc = fma(1e-9, xv[u], acc);
            __syncthreads();
            v[threadIdx.x] = acc;
            z_barrier();
        } else if (which == 16) {
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            double a = warp_sum(acc + threadIdx.x);
            double* sm = reinterpret_cast<double*>(dsm) + (it & 1) * 64;
            if (lane == 0) sm[w] = a;
            __syncthreads();
            double t = 0.0;
            for (int i = 0; i < kZT / 32; ++i) t += sm[i];
            acc += t;
            z_barrier();
        } else {
            v[threadIdx.x] = acc;
            asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
        }
    }
    if (G.rank == 0 && threadIdx.x == 0) cycles_out[0] = clock64() - t0;
    if (acc == 123.456) gbuf[0] = acc;
    z_barrier();
}
#endif

int z_log2_lanes(double avg, int rows, int trips = 1) {
    int t = 1;
    while (t < 32 && (double)t * 4.0 < avg) t <<= 1;            // up to 4 entries per lane: one batch of gathers in flight
    while (t > 1 && (int64_t)rows * t > (int64_t)trips * kZT) t >>= 1;   // ... but every row of the slice in `trips` trips if possible
    int lt = 0;
    while ((1 << lt) < t) ++lt;
    return lt;
}

void z_gen(std::vector<int>& prog, int k, int kd, int J, bool wcycle, bool zero) {
    if (k == kd) { prog.push_back(Z_LEAF | (k << 8) | ((zero ? 1 : 0) << 16)); return; }
    prog.push_back(Z_PRE | (k << 8) | ((zero ? 1 : 0) << 16));
    prog.push_back(Z_RESTRICT | (k << 8));
    z_gen(prog, k + 1, kd, J, wcycle, true);                                  // MG_Wcycle.m:28
    if (wcycle && (k + 1 != J - 1)) z_gen(prog, k + 1, kd, J, wcycle, false);    // :30 (the coarsest solve ignores its guess)
    prog.push_back(Z_PROLONG | (k << 8));
    prog.push_back(Z_POST | (k << 8));
}

}  // namespace

// Launches the kernel when the hierarchy qualifies (returns false otherwise, nothing launched): a dense tail from level
// kd >= 1, smoothing on, at most kZMaxL levels, the bigraph smoother on level 0 only, all vectors within the shared
// memory of 16 CTAs.  hist: 2*hl doubles (relk | rho), iout: 4 ints, as persist_solve reads them back.
bool dsm_cluster_solve(ssn_ctx* c, Hierarchy& H, const double* b, double* x, const AmgOptions& o, bool wcycle, double* hist, int hl,
                       int* iout, const ssn_pcg_options* leaf_pcg) {
    // leaf_pcg: the two-level method (twogrid_bigph): the coarsest level is solved by PCG with these options inside the kernel
    const int J = H.J, kd = leaf_pcg ? J - 1 : H.dense_from;
    if (kd < 1 || kd >= J || kd + 1 > kZMaxL || H.smoth < 1) return false;
    if (!leaf_pcg && (H.lv[kd].N > kZXs || H.lv[kd].B.p == nullptr)) return false;
    if (leaf_pcg && (leaf_pcg->precd != 2 || leaf_pcg->guess_dev != nullptr || leaf_pcg->maxit < 0)) return false;
    for (int k = 1; k <= kd; ++k) if (H.lv[k].bigph) return false;
    const int ncta = kZCta;
    ZArgs a{};
    // ---- layout of a CTA's shared memory.  Pass 1: without halos, to see what is left; pass 2: with the halo capacities.
    // SSN_DSM_HALO=1 (opt-in): measured on a B200 at the benchmarked state the halo copies LOSE (4.49 ms against 4.17 per solve):
    // a CTA references only ~30 distinct remote elements per level there, and the pull adds a remote round trip and a CTA
    // barrier to every sweep.  What set the time of a sweep was the tail of the longest row (z_row_dot), not the remote gathers.
    int use_halo = 0;
    { const char* e = getenv("SSN_DSM_HALO"); if (e && e[0] == '1') use_halo = 1; }
    int hcap[kZMaxL] = {0};
    size_t off = 0;
    for (int pass = 0; pass < 2; ++pass) {
        off = kOffVec;
        for (int k = 0; k <= kd; ++k) {
            Level& L = H.lv[k];
            ZLevel& z = a.lv[k];
            z.N = L.N; z.bigph = (k == 0 && L.bigph) ? 1 : 0;
            z.Nf = z.bigph ? L.Nf : L.N;
            if (z.bigph && (z.Nf <= 0 || z.Nf >= z.N)) return false;
            z.rpf = (z.Nf + ncta - 1) / ncta;
            z.rpc = (z.N - z.Nf + ncta - 1) / ncta;
            const int nl = z.rpf + z.rpc;
            const int sz0 = ((nl * 8 + 15) / 16) * 16;
            const int szh = (((nl + hcap[k]) * 8 + 15) / 16) * 16;
            if ((size_t)szh >= ((size_t)1 << 23)) return false;
            int nvec = 4;                                       // E, R, G, ALT
            z.xslot = 0;
            if (k == 0) { z.xslot = z.bigph ? ZV_ALT : 4; nvec = z.bigph ? 4 : 5; }     // the in-place smoother needs no ALT copy
            if (k == kd) nvec = leaf_pcg ? 5 : 3;               // PCG leaf: E (solution), R (right-hand side), G (q / w), ALT (p), 4 (residual)
            // the slots that A-gathers read (E, the ping-pong copy / x) carry a halo area behind the CTA's slice
            for (int sl_ = 0; sl_ < 5; ++sl_) z.so[sl_] = 0;
            const int order[5] = {ZV_E, ZV_ALT, 4, ZV_R, ZV_G};
            for (int q = 0; q < 5; ++q) {
                const int sl_ = order[q];
                if (sl_ >= nvec) continue;                       // slots of this level: 0 .. nvec-1
                const bool gathered = (sl_ == ZV_E || sl_ == ZV_ALT || sl_ == 4);
                z.so[sl_] = (int)off;
                off += (size_t)(gathered ? szh : sz0);
            }
            z.hb = sz0; z.hcap = hcap[k]; z.hn = 0;
            z.hoff = (int)off;
            off += (size_t)(((hcap[k] * 4 + 15) / 16) * 16);
            z.poff = (int)off;
            z.rsbytes = (((nl + 2) * 4 + 15) / 16) * 16;
            off += (size_t)z.rsbytes + ((k < kd || leaf_pcg) ? (size_t)2 * (((nl * 8 + 15) / 16) * 16) : 0);
            const double avgA = L.N ? (double)L.A.nnz / L.N : 0.0;
            z.ltA = z_log2_lanes(avgA, nl, kZSlots);
            z.ltG = z_log2_lanes(avgA, std::max(z.rpf, z.rpc), kZSlots);
            z.A = ZMat{L.A.ptr.p, L.A.idx.p, L.A.val.p, nullptr};
            z.Pu = ZMat{nullptr, nullptr, nullptr, nullptr}; z.Td = z.Pu;
            if (k < kd) {
                Level& Lc = H.lv[k + 1];
                z.Pu = ZMat{Lc.P.ptr.p, Lc.P.idx.p, Lc.P.val.p, nullptr};
                z.ltP = z_log2_lanes(L.N ? (double)Lc.P.nnz / L.N : 0.0, nl);
            }
            if (k >= 1) {
                z.Td = ZMat{L.Pt.ptr.p, L.Pt.idx.p, L.Pt.val.p, nullptr};
                z.ltT = z_log2_lanes(L.N ? (double)L.Pt.nnz / L.N : 0.0, nl);
            }
            z.dinv = L.dinv.p; z.Axi = L.Axi.p; z.xx = L.xx; z.B = (k == kd && !leaf_pcg) ? L.B.p : nullptr;
        }
        if (pass == 0) {
            if (off > (size_t)c->smem_optin - 1024) return false;
            if (!use_halo) break;
            // what is left goes to the halos: up to 1024 entries per level, scaled down together when that is too much
            const double left = (double)((size_t)c->smem_optin - 1024 - off) - 64.0 * (kd + 1);
            double need = 0.0; int want[kZMaxL];
            for (int k = 0; k <= kd; ++k) {
                const ZLevel& z = a.lv[k];
                const int ng = (k == kd) ? 1 : ((k == 0 && !z.bigph) ? 3 : 2);
                want[k] = std::max(0, std::min(1024, z.N - (z.rpf + z.rpc)));
                need += (double)want[k] * (8.0 * ng + 4.0);
            }
            const double fsc = (need > 0.0 && left > 0.0) ? std::min(1.0, left / need) : 0.0;
            bool any = false;
            for (int k = 0; k <= kd; ++k) { hcap[k] = ((int)(want[k] * fsc)) & ~1; any = any || hcap[k] > 0; }
            if (!any) break;
        }
    }
    if (leaf_pcg) {
        // what is left of the shared memory holds the long rows' tails of the PCG leaf (12 bytes per entry); at most every entry of
        // a CTA's slice past the first 8 of each row would be asked for
        off = (off + 15) / 16 * 16;
        const size_t left = ((size_t)c->smem_optin - 1024 > off) ? (size_t)c->smem_optin - 1024 - off : 0;
        int64_t cap = (int64_t)(left / 12) & ~(int64_t)3;
        const int64_t most = (H.lv[kd].A.nnz + 3) & ~(int64_t)3;
        if (cap > most) cap = most;
        { const char* e = getenv("SSN_PCG_TAIL"); if (e && e[0] == '0') cap = 0; }
        { const char* e = getenv("SSN_PCG_LT0"); if (e && e[0] == '1') a.lv[kd].ltA = 0; }   // tests: one lane per row whatever the row lengths
        a.tail_off = (int)off; a.tail_cap = (int)cap;
        off += (size_t)cap * 12;
    }
    const size_t smem = off;
    if (smem > (size_t)c->smem_optin - 1024) return false;
    std::vector<int> prog;
    z_gen(prog, 0, kd, J, wcycle, true);
    if ((int)prog.size() > kZProgMax) return false;
    if (leaf_pcg) for (int& op : prog) if ((op & 0xff) == Z_LEAF) { if (!((op >> 16) & 1)) return false; op = (op & ~0xff) | Z_PCG; }
#ifndef SSN_EMU
    static int ok16 = -1;                                   // can the device co-schedule a 16-CTA cluster of this kernel? (probed once)
    if (ok16 < 0) {
        ok16 = 0;
        if (cudaFuncSetAttribute(dsm_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin - 1024) == cudaSuccess &&
            cudaFuncSetAttribute(dsm_solve_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(kZT); cfg.dynamicSmemBytes = (size_t)c->smem_optin - 1024; cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, dsm_solve_kernel, &cfg) == cudaSuccess && nclusters >= 1) ok16 = 1;
        }
        (void)cudaGetLastError();
    }
    if (!ok16) return false;
#endif
    // gather tables (one uint32 per matrix entry) and the op list: scratch of this launch
    size_t nloc = 0;
    for (int k = 0; k <= kd; ++k) {
        nloc += (size_t)H.lv[k].A.nnz;
        if (k < kd) nloc += (size_t)H.lv[k + 1].P.nnz;
        if (k >= 1) nloc += (size_t)H.lv[k].Pt.nnz;
    }
    Buf<uint32_t> loc(c, nloc + 1);
    Buf<int> dprog(c, prog.size());
    size_t at_loc = 0;
    for (int k = 0; k <= kd; ++k) {
        a.lv[k].A.loc = loc.p + at_loc; at_loc += (size_t)H.lv[k].A.nnz;
        if (k < kd) { a.lv[k].Pu.loc = loc.p + at_loc; at_loc += (size_t)H.lv[k + 1].P.nnz; }
        if (k >= 1) { a.lv[k].Td.loc = loc.p + at_loc; at_loc += (size_t)H.lv[k].Pt.nnz; }
    }
    upload_small(c, dprog.p, prog.data(), sizeof(int) * prog.size());
    a.kd = kd; a.smoth = H.smoth; a.isnsp = o.isnsp; a.maxit = o.maxit;
    a.b = b; a.x = x; a.retol = o.retol;
    a.relk = hist; a.rho = hist + hl; a.it_out = iout;
    a.prog = dprog.p; a.nprog = (int)prog.size();
    if (leaf_pcg) {
        const double retol = (!(leaf_pcg->retol < 0) && leaf_pcg->retol == leaf_pcg->retol) ? leaf_pcg->retol : 1e-11;   // PCG.m: [] -> 1e-11
        a.leaf_pcg = 1; a.pcg_maxit = leaf_pcg->maxit; a.pcg_tol2 = retol * retol;
    }
    { const char* e = getenv("SSN_DSM_NOREG"); a.noreg = (e && e[0] == '1') ? 1 : 0; }
    Phase ph(c, "solve.dsm_solve_kernel");
#ifdef SSN_EMU
    emu_launch_cluster(c, dsm_solve_kernel, ncta, kZT, smem, a);
#else
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(kZT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        KernelTimer kt(c);
        SSN_CUDA(cudaLaunchKernelEx(&cfg, dsm_solve_kernel, a));
        c->launches++;
    }
    // the scratch is stream-ordered pool memory: freed behind the kernel on the same stream
#endif
    return true;
}

#ifndef SSN_EMU
// cycles per iteration of one of the building blocks above (which = 10..17; ng gathers per thread rides in which / 100)
double dsm_bench(ssn_ctx* c, int iters, int which_ng) {
    const int which = which_ng % 100, ng = std::max(1, std::min(16, which_ng / 100));
    Buf<double> buf(c, 16 * 1024); buf.zero();
    Buf<long long> out(c, 1);
    const size_t smem = 4096 + 1024 * 8;
    if (cudaFuncSetAttribute(dsm_bench_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { (void)cudaGetLastError(); return -1.0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kZCta); cfg.blockDim = dim3(kZT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kZCta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, dsm_bench_kernel, buf.p, iters, which, ng, out.p) != cudaSuccess) { (void)cudaGetLastError(); return -1.0; }
    return (double)read_scalar(c, out.p) / (double)iters;
}
#endif

}  // namespace ssn
