// amg.cuh -- the Class_AMG hierarchy (the reference's globals Ack/Prok/Rk/J/smoth_it,
// AMG/Class_AMG.m:42-43) as a context-owned handle, plus the internal AMG interfaces.
#pragma once
#include "common.cuh"
#include "sparse.cuh"

namespace ssn {

// Device-visible description of one level (array uploaded after setup).
struct LevelDev {
    int N, Nf, bigph;                                  // bigph: level-1 block Gauss-Seidel smoother
    const int* ap; const int* ai; const double* av;    // A_k
    const int* pp; const int* pi; const double* pv;    // Pro_k   (N_{k-1} x N_k), null at level 1
    const int* tp; const int* ti; const double* tv;    // Pro_k'  (N_k x N_{k-1})
    const double* dinv;   // bigph: 1/diag(A) (invV ; invT, Class_AMG.m:56-57); else 0.5*(1/diag) (:72,:84)
    const double* Axi;    // A*ones
    double xx;            // ones'*A*ones
    double* r; double* e; double* g;                   // work vectors
    double* pcg;                                       // 5*N scratch (small levels)
    const double* B;                                   // dense cycle operator (tail levels), else null
};

struct Level {
    int N = 0, Nf = 0, bigph = 0;
    Csr A, P, Pt;
    Buf<double> dinv, Axi, r, e, g, pcg;
    double xx = 0.0;
    bool xx_known = false;         // xx was computed with the level (fused setup kernel); else amg_setup reads it back
    // setup trace (kept for parity tests: C/F split and strength flags that produced level k+1)
    Buf<uint8_t> isC;
    // dense cycle operator of this level (tail levels only; amg_solve.cu "dense tail")
    Buf<double> B;
};

constexpr int kClusterSize = 8;
constexpr int kClusterLevels = 12;
struct ClusterPlan {
    int nl;
    int staged[kClusterLevels];
    int smem_off[kClusterLevels];              // byte offset of the staged slice in dynamic smem
    int rpc[kClusterLevels];                   // rows per CTA = ceil(N / cluster size)
    int cap[kClusterLevels];                   // capacity (entries) of the staged slice
};

// ---- fused coarsening of the small levels (amg_setup_fused.cu)
constexpr int kFusedMaxN = 4096;               // levels with at most this many rows ...
constexpr int64_t kFusedMaxNnz = 1 << 18;      // ... and nonzeros are coarsened by one kernel
constexpr int kFusedMaxLevels = 24;
constexpr int kFusedOverflow = 1;              // status: out of arena / a product row wider than the accumulator
struct FusedLevel {                            // one level built by the kernel: arrays inside the arena
    int N, nnzA, nnzP, parentN; double xx;
    int* ap; int* ai; double* av; int* pp; int* pi; double* pv; int* tp; int* ti; double* tv;
    double* dinv; double* Axi; uint8_t* parent_isC;
};

struct Hierarchy {
    std::vector<Buf<unsigned char>> arenas;    // memory of the levels built by the fused kernel (declared first: freed last)
    std::vector<Level> lv;
    int J = 0;
    int smoth = 0;
    int small_from = 0;            // levels >= small_from are solved by the single-block kernel
    int cluster_from = 1 << 30;    // levels >= cluster_from are solved by the 8-CTA cluster kernel
    int dense_from = 1 << 30;      // levels >= dense_from are applied as dense cycle operators B_k
    int dense_isnsp = -1, dense_w = -1;   // the (isnsp, cycle) the operators were built for
    ClusterPlan cluster_plan{};
    size_t cluster_smem = 0;
    Buf<LevelDev> dev;
    std::vector<LevelDev> hdev;    // host copy of dev
    Buf<double> part;              // reduction partials (multi-block kernels)
    Buf<double> scal;              // small device scalars
};

struct AmgOptions {                // amg_options with the isempty() defaults applied
    double retol; int bigph; int maxit; double theta; int smoth; int cycle; int isnsp; int inter; int fnode;
    const double* guess;
};
AmgOptions resolve_options(const ssn_amg_options* o);

// ---- setup (amg_setup.cu)
void rng_reset(ssn_ctx* c, uint32_t seed);
void rng_rand(ssn_ctx* c, int64_t count, double* out_dev);
Csr strength_matrix(ssn_ctx* c, const CsrView& A, int which);
// strength flags aligned with A's entries: as[e] = (S(i,j) >= theta)
void strength_flags(ssn_ctx* c, const CsrView& A, double theta, uint8_t* as_flags);
void mis_set(ssn_ctx* c, const CsrView& A, double theta, uint8_t* isC, uint8_t* isF, Buf<uint8_t>& as_flags);
void cf_split(ssn_ctx* c, const CsrView& S, uint8_t* indC, uint8_t* indF);
Csr flags_to_csr(ssn_ctx* c, const CsrView& A, const uint8_t* as_flags);
// one coarsening step: returns Ac, Pro (and keeps isC / strength flags if requested)
void transfer(ssn_ctx* c, const CsrView& A, const AmgOptions& o, int level_J, Csr& Ac, Csr& Pro,
              Buf<uint8_t>* isC_out, Buf<uint8_t>* as_out, Csr* Pt_out = nullptr);
// max_levels > 0 stops the coarsening after that many levels (twogrid_bigph builds exactly two)
void amg_setup(ssn_ctx* c, const CsrView& A, const AmgOptions& o, int max_levels = 0);
void amg_clear(ssn_ctx* c);
// coarsens below the last level of H in one kernel while N <= kFusedMaxN; false: not applicable, H and the random stream untouched
bool fused_small_levels(ssn_ctx* c, Hierarchy& H, const AmgOptions& o, int thr, int max_new_levels);
int coarsest_threshold(int64_t N);

// ---- solve (amg_solve.cu)
void mg_cycle(ssn_ctx* c, const double* r_dev, int isnsp, int k1, double* e_dev, bool wcycle, bool e_is_zero);
void class_amg(ssn_ctx* c, const CsrView& A, const double* b, const AmgOptions& o, bool keep, double* x, int* it_out,
               double* rel_res_out, double* rel_resk, double* rhok, int* hist_len);
// [x,it,rel_res,rel_resk,rhok] = twogrid_bigph(A,b,amg_options) -- AMG/twogrid_bigph.m; generic: AMG/twogrid.m
void twogrid_bigph(ssn_ctx* c, const CsrView& A, const double* b, const AmgOptions& o, double* x, int* it_out,
                   double* rel_res_out, double* rel_resk, double* rhok, int* hist_len, bool generic = false);
void pcg_solve(ssn_ctx* c, const CsrView& H, const double* e, const ssn_pcg_options* opts, double* d, int* it_out,
               double* res_out, double* resk_host);

// ---- triangular preconditioner setup and Jk assembly on the device (trifactor.cu)
struct TriFactors {
    Buf<int> lp, li, up, ui, lrows, llev, urows, ulev; Buf<double> lv, uv, mid;
    int nl = 0, nu = 0; bool has_mid = false;
};
void build_tri_factors_device(ssn_ctx* c, const CsrView& H, int precd, TriFactors& F);
// Jk = bk1*speye(m+n) + (T + H0)/tk -- Class1/APD_SsN_Class1.m:147,151
void jk_system(ssn_ctx* c, const ssn_prob_data* pd, Csr& Jk);

void debug_cycles(unsigned long long* out64, bool reset);
void debug_cycles_persist(unsigned long long* out256, bool reset);
void debug_cycles_dsm(unsigned long long* out256, bool reset);
double barrier_bench(ssn_ctx* c, int iters, int which);
double dsm_bench(ssn_ctx* c, int iters, int which_ng);
void build_cluster_plan(ssn_ctx* c, Hierarchy& H);
// Class_AMG's solve loop in one 16-CTA cluster with the level vectors in distributed shared memory (amg_cluster.cu);
// false: the hierarchy does not qualify, nothing was launched
bool dsm_cluster_solve(ssn_ctx* c, Hierarchy& H, const double* b, double* x, const AmgOptions& o, bool wcycle, double* hist, int hl,
                       int* iout, const ssn_pcg_options* leaf_pcg = nullptr);

// ---- dispatch (solvers.cu)
void components(ssn_ctx* c, const CsrView& A, int* blocks, int* sizes, int* perm, int* r, int* ncomp);
void rescaled_system(ssn_ctx* c, const ssn_prob_data* pd, Csr& Ae, double* f, Buf<double>& qp, Buf<double>& Kd);
// twogrid = false: Hybrid_AMG.m (Class_AMG per component); true: Hybrid_twogrid.m (twogrid_bigph per component)
void hybrid_amg(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, double* zeta, int* itamg,
                double* resamg, int* info, bool twogrid = false);
void aug_pcg(ssn_ctx* c, const ssn_prob_data* pd, const ssn_pcg_options* opts, double* zeta, int* itpcg,
             double* respcg, int* info);

}  // namespace ssn
