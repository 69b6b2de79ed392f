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
constexpr int kZT = 1024;
typedef uint32_t zaddr;
#endif
constexpr int kZMaxL = 10;                      // explicit levels + the dense leaf
constexpr int kZCta = 16;
constexpr int kZProgMax = 1024;
constexpr int kZXs = 2048;                      // largest dense leaf

enum { Z_PRE = 0, Z_RESTRICT = 1, Z_LEAF = 2, Z_PROLONG = 3, Z_POST = 4 };
enum { ZV_E = 0, ZV_R = 1, ZV_G = 2, ZV_ALT = 3 };          // vector slots of a level (x: ZLevel::xslot, level 0 only)

struct ZMat { const int* rp; const int* ci; const double* cv; uint32_t* loc; };
struct ZLevel {
    int N, Nf;                                  // Nf: rows of the first segment (bigraph level: fnode; else N)
    int rpf, rpc;                               // rows per CTA of the two segments
    int voff, stride, xslot, bigph;             // byte offset of the level's vectors in a CTA's shared memory, bytes per vector
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
};

constexpr int kOffSlots = 1024, kOffSumR = 1536, kOffDot = 1616, kOffCur = 1696, kOffLv = 2048, kOffProg = 4096,
              kOffXs = 8192, kOffVec = kOffXs + kZXs * 8;
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
// owner CTA (bits 24..) and byte offset inside the owner's slice (bits 0..23) of element j of a vector of level L
__device__ __forceinline__ uint32_t z_loc(const ZLevel& L, int j) {
    int owner, l;
    if (j < L.Nf) { owner = j / L.rpf; l = j - owner * L.rpf; }
    else { const int jj = j - L.Nf; owner = jj / L.rpc; l = L.rpf + jj - owner * L.rpc; }
    return ((uint32_t)owner << 24) | (uint32_t)(l * 8);
}
__device__ __forceinline__ double* z_vec(const ZTeam& G, const ZLevel& L, int slot) { return reinterpret_cast<double*>(G.dsm + L.voff + slot * L.stride); }
__device__ __forceinline__ int z_vb(const ZLevel& L, int slot) { return L.voff + slot * L.stride; }
__device__ __forceinline__ double z_gather(const ZTeam& G, int vb, uint32_t lc) { return z_ld(z_map(G.base + (zaddr)(vb + (int)(lc & 0xffffffu)), (int)(lc >> 24))); }

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

// For the local rows [l0, l1) of level L (2^lt lanes per row): s = sum over the row's entries of M of value * v[column],
// v the vector of the COLUMN level that starts `vb` bytes into every CTA's shared memory; filt = 1 / 2 keeps only the
// entries whose column lies in the second / first segment (the two halves of the bigraph level); gather = false: s = 0.
// Then epi(l, row, s) on the row's first lane.  Ends WITHOUT a barrier.
template <class Epi>
__device__ __forceinline__ void z_rows(const ZTeam& G, const ZLevel& L, const ZMat& M, int lt, int l0, int l1, int vb, int filt,
                                       int seg_bytes, bool gather, Epi&& epi) {
    const int tpr = 1 << lt, sub = threadIdx.x & (tpr - 1), rpt = kZT >> lt;
    for (int lb = l0; lb < l1; lb += rpt) {
        const int l = lb + (threadIdx.x >> lt);
        const int row = (l < l1) ? z_row(L, G.rank, l) : -1;
        double s = 0.0;
        if (row >= 0 && gather) {
            int e = M.rp[row] + sub;
            const int e1 = M.rp[row + 1];
            for (; e + 3 * tpr < e1; e += 4 * tpr) {
                uint32_t lc[4]; double v[4], xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { lc[u] = M.loc[e + u * tpr]; v[u] = M.cv[e + u * tpr]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool second = (int)(lc[u] & 0xffffffu) >= seg_bytes;
                    const bool keep = filt == 0 || (second == (filt == 1));
                    xv[u] = keep ? z_gather(G, vb, lc[u]) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) s = fma(v[u], xv[u], s);
            }
            for (; e < e1; e += tpr) {
                const uint32_t lc = M.loc[e];
                const bool second = (int)(lc & 0xffffffu) >= seg_bytes;
                const bool keep = filt == 0 || (second == (filt == 1));
                if (keep) s = fma(M.cv[e], z_gather(G, vb, lc), s);
            }
        }
        for (int o = tpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row >= 0 && sub == 0) epi(l, row, s);
    }
}

// loc tables of the rows this CTA owns (one thread per row); returns true when an off-diagonal entry joins two rows of
// the same segment (the two-half-sweep smoother needs V and T diagonal)
__device__ bool z_build_loc(const ZTeam& G, const ZLevel& Lrow, const ZLevel& Lcol, const ZMat& M, bool check) {
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

}  // namespace

#ifdef SSN_EMU
void dsm_solve_kernel(const ZArgs a) {
    unsigned char* dsm = emu::cluster_smem[blockIdx.x];
#else
__global__ void __launch_bounds__(kZT, 1) dsm_solve_kernel(const ZArgs a) {
    extern __shared__ __align__(16) unsigned char dsm[];
#endif
    ZTeam G{z_rank(), z_ncta(), 0, dsm, z_local(dsm)};
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
    const bool lead = (G.rank == 0 && threadIdx.x == 0);
    const int smoth = a.smoth;
    const bool nsp = a.isnsp != 0;

    // r = b - A*x ; sum(r), sum(r^2)                                    Class_AMG.m:89 / :96,:102
    auto outer_residual = [&](double& s1, double& s2) {
        s1 = 0.0; s2 = 0.0;
        z_rows(G, L0, L0.A, L0.ltA, 0, nl0, z_vb(L0, L0.xslot), 0, 0, true, [&](int l, int row, double s) {
            const double ri = a.b[row] - s; r0[l] = ri; s1 += ri; s2 = fma(ri, ri, s2);
        });
        z_sum2(G, s1, s2);
    };
    // smoth damped-Jacobi sweeps on level k (ping-pong E <-> ALT); returns Axi'e of the result        MG_Wcycle.m:15-23
    auto jacobi = [&](int k, bool ez, double sr, double dotAe) -> double {
        const ZLevel& L = sl[k];
        const int nl = L.rpf + L.rpc;
        const double* r = z_vec(G, L, ZV_R);
        int cur = st_cur[k];
        for (int s = 0; s < smoth; ++s) {
            const double coef = nsp ? (sr - dotAe) / L.xx : 0.0;
            const double* ec = z_vec(G, L, cur ? ZV_ALT : ZV_E);
            double* ea = z_vec(G, L, cur ? ZV_E : ZV_ALT);
            double part = 0.0, dummy = 0.0;
            z_rows(G, L, L.A, L.ltA, 0, nl, z_vb(L, cur ? ZV_ALT : ZV_E), 0, 0, !ez, [&](int l, int row, double d) {
                const double axi = L.Axi[row], di = L.dinv[row], ei = ez ? 0.0 : ec[l];
                const double en = ei + coef + di * ((r[l] - d) - axi * coef);
                ea[l] = en;
                part = fma(axi, en, part);
            });
            z_sum2(G, part, dummy);
            dotAe = part; cur ^= 1; ez = false;
        }
        st_cur[k] = cur;
        return dotAe;
    };
    // smoth block Gauss-Seidel sweeps on the bigraph level 0, in place in E; returns Axi'e              Class_AMG.m:48-59
    auto gauss_seidel = [&](bool post, bool ez, double sr, double dotAe) -> double {
        const ZLevel& L = L0;
        const double* r = r0;
        double* e = z_vec(G, L, ZV_E);
        const int vbe = z_vb(L, ZV_E), segb = L.rpf * 8;
        // pre: first-segment rows with the kernel correction, then second-segment rows; post: the other way round
        const int a0 = post ? L.rpf : 0, a1 = post ? L.rpf + L.rpc : L.rpf;
        const int b0 = post ? 0 : L.rpf, b1 = post ? L.rpf : L.rpf + L.rpc;
        for (int s = 0; s < smoth; ++s) {
            const double coef = nsp ? (sr - dotAe) / L.xx : 0.0;
            double part = 0.0, dummy = 0.0;
            z_rows(G, L, L.A, L.ltG, a0, a1, vbe, post ? 2 : 1, segb, !ez, [&](int l, int row, double d) {
                const double axi = L.Axi[row];
                const double en = coef + L.dinv[row] * ((r[l] - d) - axi * coef);
                e[l] = en;
                part = fma(axi, en, part);
            });
            z_barrier();
            z_rows(G, L, L.A, L.ltG, b0, b1, vbe, post ? 1 : 2, segb, true, [&](int l, int row, double d) {
                const double en = L.dinv[row] * (r[l] - d);
                e[l] = en;
                part = fma(L.Axi[row], en, part);
            });
            z_sum2(G, part, dummy);
            dotAe = part; ez = false;
        }
        return dotAe;
    };

    double s1, s2;
    outer_residual(s1, s2);
    const double res0 = sqrt(s2);
    double sum_r0 = s1;
    double res_prev = res0, rel_prev = 1.0;
    int it = 0, hist = 1;
    if (lead) { a.relk[0] = 1.0; a.rho[0] = NAN; a.it_out[2] = 0; }
    if (res0 == 0.0) {
        if (lead) { a.relk[0] = 0.0; a.rho[0] = INFINITY; a.it_out[0] = 0; a.it_out[1] = 1; }
        z_barrier();
        return;
    }
    it = 1;
    while (rel_prev > a.retol && it <= a.maxit) {                       // Class_AMG.m:95
        st_sum_r[0] = sum_r0;
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
                const double d = (k == 0 && L.bigph) ? gauss_seidel(post, op == Z_PRE && zero, sr, d0)
                                                     : jacobi(k, op == Z_PRE && zero, sr, d0);
                st_dot[k] = d;
            } else if (op == Z_RESTRICT) {                              // r_{k+1} = Pro' (r - A e)            MG_Wcycle.m:26
                const double* r = z_vec(G, L, ZV_R);
                double* g = z_vec(G, L, ZV_G);
                const int es = (k == 0 && L.bigph) ? ZV_E : (st_cur[k] ? ZV_ALT : ZV_E);
                z_rows(G, L, L.A, L.ltA, 0, nl, z_vb(L, es), 0, 0, true, [&](int l, int, double d) { g[l] = r[l] - d; });
                z_barrier();
                const ZLevel& Lc = sl[k + 1];
                double* rc = z_vec(G, Lc, ZV_R);
                double sy = 0.0, dummy = 0.0;
                z_rows(G, Lc, Lc.Td, Lc.ltT, 0, Lc.rpf + Lc.rpc, z_vb(L, ZV_G), 0, 0, true, [&](int l, int, double d) { rc[l] = d; sy += d; });
                z_sum2(G, sy, dummy);
                st_sum_r[k + 1] = sy;
            } else if (op == Z_PROLONG) {                               // e += Pro e_{k+1}, with Axi'e        MG_Wcycle.m:32
                const ZLevel& Lc = sl[k + 1];
                const int cs = (k + 1 == kd) ? ZV_E : (st_cur[k + 1] ? ZV_ALT : ZV_E);
                const int es = (k == 0 && L.bigph) ? ZV_E : (st_cur[k] ? ZV_ALT : ZV_E);
                double* e = z_vec(G, L, es);
                double swy = 0.0, dummy = 0.0;
                z_rows(G, L, L.Pu, L.ltP, 0, nl, z_vb(Lc, cs), 0, 0, true, [&](int l, int row, double d) {
                    const double v = e[l] + d; e[l] = v; swy = fma(L.Axi[row], v, swy);
                });
                z_sum2(G, swy, dummy);
                st_dot[k] = swy;
            } else {                                                    // Z_LEAF: e = B r, or e += B (r - A e)
                const double* r = z_vec(G, L, ZV_R);
                double* e = z_vec(G, L, ZV_E);
                int in_slot = ZV_R;
                if (!zero) {
                    double* g = z_vec(G, L, ZV_G);
                    z_rows(G, L, L.A, L.ltA, 0, nl, z_vb(L, ZV_E), 0, 0, true, [&](int l, int, double d) { g[l] = r[l] - d; });
                    z_barrier();
                    in_slot = ZV_G;
                }
                const int n = L.N, vb = z_vb(L, in_slot);
                for (int j = threadIdx.x; j < n; j += kZT) xs[j] = z_gather(G, vb, z_loc(L, j));
                __syncthreads();
                const int lane = threadIdx.x & 31;
                for (int l = threadIdx.x >> 5; l < nl; l += kZT / 32) {
                    const int row = z_row(L, G.rank, l);
                    if (row < 0) continue;
                    const double* Br = L.B + (size_t)row * n;
                    double acc[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int j0 = 0; j0 < n; j0 += 256) {
                        double bv[8], xv[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) { const int j = j0 + u * 32 + lane; bv[u] = (j < n) ? Br[j] : 0.0; xv[u] = (j < n) ? xs[j] : 0.0; }
#pragma unroll
                        for (int u = 0; u < 8; ++u) acc[u & 3] = fma(bv[u], xv[u], acc[u & 3]);
                    }
                    const double s = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
                    if (lane == 0) e[l] = zero ? s : (e[l] + s);
                }
                z_barrier();
            }
        }
        // ---------------- x += e ; r = b - A*x ; res = norm(r)          Class_AMG.m:96-104
        {
            const double* et = z_vec(G, L0, L0.bigph ? ZV_E : (st_cur[0] ? ZV_ALT : ZV_E));
            for (int l = threadIdx.x; l < nl0; l += kZT) x0[l] += et[l];
            z_barrier();
        }
        outer_residual(s1, s2);
        sum_r0 = s1;
        const double res = sqrt(s2);
        const double rel_res = res / res0, rho = res / res_prev;
        if (lead) { a.relk[it] = rel_res; a.rho[it] = rho; }
        res_prev = res; rel_prev = rel_res;
        ++it; ++hist;
        if (rho > 1.0) break;                                           // Class_AMG.m:106
    }
    for (int l = threadIdx.x; l < nl0; l += kZT) { const int row = z_row(L0, G.rank, l); if (row >= 0) a.x[row] = x0[l]; }
    if (lead) { a.it_out[0] = it - 1; a.it_out[1] = hist; }
    z_barrier();                                            // no CTA exits while a peer may still read its shared memory
}

namespace {

int z_log2_lanes(double avg, int rows) {
    int t = 1;
    while (t < 32 && (double)t * 4.0 < avg) t <<= 1;            // up to 4 entries per lane: one batch of gathers in flight
    while (t > 1 && (int64_t)rows * t > kZT) t >>= 1;           // ... but every row of the slice in one trip if possible
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
                       int* iout) {
    const int kd = H.dense_from, J = H.J;
    if (kd < 1 || kd >= J || kd + 1 > kZMaxL || H.smoth < 1) return false;
    if (H.lv[kd].N > kZXs || H.lv[kd].B.p == nullptr) return false;
    for (int k = 1; k <= kd; ++k) if (H.lv[k].bigph) return false;
    const int ncta = kZCta;
    ZArgs a{};
    size_t off = kOffVec;
    for (int k = 0; k <= kd; ++k) {
        Level& L = H.lv[k];
        ZLevel& z = a.lv[k];
        z.N = L.N; z.bigph = (k == 0 && L.bigph) ? 1 : 0;
        z.Nf = z.bigph ? L.Nf : L.N;
        if (z.bigph && (z.Nf <= 0 || z.Nf >= z.N)) return false;
        z.rpf = (z.Nf + ncta - 1) / ncta;
        z.rpc = (z.N - z.Nf + ncta - 1) / ncta;
        const int nl = z.rpf + z.rpc;
        z.stride = ((nl * 8 + 15) / 16) * 16;
        if ((size_t)z.stride >= ((size_t)1 << 24)) return false;
        z.voff = (int)off;
        int nvec = 4;                                       // E, R, G, ALT
        if (k == 0) { z.xslot = z.bigph ? ZV_ALT : 4; nvec = z.bigph ? 4 : 5; }     // the in-place smoother needs no ALT copy
        if (k == kd) nvec = 3;
        off += (size_t)nvec * z.stride;
        const double avgA = L.N ? (double)L.A.nnz / L.N : 0.0;
        z.ltA = z_log2_lanes(avgA, nl);
        z.ltG = z_log2_lanes(avgA, std::max(z.rpf, z.rpc));
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
        z.dinv = L.dinv.p; z.Axi = L.Axi.p; z.xx = L.xx; z.B = (k == kd) ? L.B.p : nullptr;
    }
    const size_t smem = off;
    if (smem > (size_t)c->smem_optin - 1024) return false;
    std::vector<int> prog;
    z_gen(prog, 0, kd, J, wcycle, true);
    if ((int)prog.size() > kZProgMax) return false;
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
    SSN_CUDA(cudaMemcpyAsync(dprog.p, prog.data(), sizeof(int) * prog.size(), cudaMemcpyHostToDevice, c->stream));
    a.kd = kd; a.smoth = H.smoth; a.isnsp = o.isnsp; a.maxit = o.maxit;
    a.b = b; a.x = x; a.retol = o.retol;
    a.relk = hist; a.rho = hist + hl; a.it_out = iout;
    a.prog = dprog.p; a.nprog = (int)prog.size();
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

}  // namespace ssn
