// trifactor.cu -- device-side setup of the triangular preconditioners of PCG.m (precd 3: SSOR, :39-44,96-99;
// precd 4: ichol, :45-51,100-101) and of the assembled KKT matrix Jk of Class1/APD_SsN_Class1.m:147,151.
//
// The factors Lf (CSR, columns ascending, diagonal LAST in its row), Uf (CSR, columns ascending, diagonal
// FIRST), their dependency levels and the rows grouped by level are what pcg_kernel's level-scheduled
// triangular solves read (amg_solve.cu).  Arithmetic follows the host construction entry by entry:
//   SSOR  Lf = D + w*L, mid = D, Uf = (w*(2-w))*(D + w*U) with w = 1.5, one rounding per product;
//   IC(0) row i in column order, L_ij = (a_ij - sum_{t<j} L_it*L_jt)/L_jj, L_ii = sqrt(a_ii - sum_t L_it^2),
//         the sum taken over the common columns in ascending order, multiply then subtract (no FMA).
// Rows of one dependency level of Lf only read rows of lower levels, so the factorisation runs level by level
// (one launch per level, a thread per row: the order inside a row is sequential by definition).
//
// The default (ssn_ctx::device_setup; run against the host construction and the oracle on a B200 by
// tests/test_zz_device_setup.py); SSN_DEVICE_SETUP=0 / ssn_set_device_setup(ctx, 0) selects the host construction.
#include "amg.cuh"
#include "sparse.cuh"

namespace ssn {

namespace {

// per row: entries below the diagonal, above it, and whether the diagonal is stored
__global__ void tri_count_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx, int* __restrict__ lcnt,
                                 int* __restrict__ ucnt, int* __restrict__ nodiag_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = 0; bool has = false;
    for (int e = ptr[i]; e < ptr[i + 1]; ++e) { const int j = idx[e]; lo += (j < i); hi += (j > i); has |= (j == i); }
    lcnt[i] = lo + 1; ucnt[i] = hi + 1;
    if (!has) *nodiag_flag = 1;
}

// SSOR: Lf = D + w*L (diagonal last), Uf = sc*(D + w*U) (diagonal first), mid = D; a missing diagonal is a stored 0
__global__ void ssor_fill_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                                 double om, double sc, const int* __restrict__ lp, int* __restrict__ li, double* __restrict__ lv,
                                 const int* __restrict__ up, int* __restrict__ ui, double* __restrict__ uv, double* __restrict__ mid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = lp[i], hi = up[i] + 1;
    double dii = 0.0;
    for (int e = ptr[i]; e < ptr[i + 1]; ++e) {
        const int j = idx[e]; const double h = val[e];
        if (j < i) { li[lo] = j; lv[lo] = __dmul_rn(om, h); ++lo; }
        else if (j > i) { ui[hi] = j; uv[hi] = __dmul_rn(sc, __dmul_rn(om, h)); ++hi; }
        else dii = h;
    }
    li[lo] = i; lv[lo] = dii;
    ui[up[i]] = i; uv[up[i]] = __dmul_rn(sc, dii);
    mid[i] = dii;
}

// IC(0): the pattern of tril(H) with the entries of H (diagonal last: columns are ascending)
__global__ void tril_fill_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                                 const int* __restrict__ lp, int* __restrict__ li, double* __restrict__ lv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = lp[i];
    for (int e = ptr[i]; e < ptr[i + 1]; ++e) { const int j = idx[e]; if (j <= i) { li[lo] = j; lv[lo] = val[e]; ++lo; } }
}

// One relaxation of lev[i] = max over the off-diagonal columns j of row i of lev[j] + 1 (0 without any).  Levels only
// grow and never pass the longest dependency path, so in-place sweeps end at it whatever the thread order.
// diag_first = 0: the diagonal is the last entry of the row (Lf); 1: the first (Uf).
__global__ void level_relax_kernel(int n, const int* __restrict__ p, const int* __restrict__ idx, int diag_first,
                                   int* lev, int* __restrict__ changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e0 = p[i] + (diag_first ? 1 : 0), e1 = p[i + 1] - (diag_first ? 0 : 1);
    int l = 0;
    for (int e = e0; e < e1; ++e) { const int v = ((volatile int*)lev)[idx[e]] + 1; l = v > l ? v : l; }
    if (l > lev[i]) { lev[i] = l; *changed = 1; }
}

__global__ void level_hist_kernel(int n, const int* __restrict__ lev, int* __restrict__ cnt, int* __restrict__ maxlev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&cnt[lev[i]], 1);
    atomicMax(maxlev, lev[i]);
}

// rows[0..count): the rows of one dependency level.  A thread per row.
__global__ void ic0_level_kernel(int count, const int* __restrict__ rows, const int* __restrict__ lp, const int* __restrict__ li,
                                 double* lv, int* __restrict__ notspd_flag) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int i = rows[t];
    const int r0 = lp[i], r1 = lp[i + 1];
    for (int e = r0; e < r1; ++e) {
        const int j = li[e];
        double sacc = lv[e];
        int a0 = r0; const int a1 = e;                                       // row i so far (columns < j)
        int b0 = (j < i) ? lp[j] : r0; const int b1 = (j < i) ? lp[j + 1] - 1 : e;   // row j without its diagonal
        while (a0 < a1 && b0 < b1) {
            const int ca = li[a0], cb = li[b0];
            if (ca == cb) { sacc = __dsub_rn(sacc, __dmul_rn(lv[a0], lv[b0])); ++a0; ++b0; }
            else if (ca < cb) ++a0; else ++b0;
        }
        if (j < i) lv[e] = __ddiv_rn(sacc, lv[lp[j + 1] - 1]);
        else { if (!(sacc > 0.0)) *notspd_flag = 1; lv[e] = __dsqrt_rn(sacc); }
    }
}

// dependency levels of a triangular CSR matrix and its rows grouped by level (rows ascending inside a level)
void levels_and_buckets(ssn_ctx* c, int n, const int* p, const int* idx, int diag_first, Buf<int>& lev, Buf<int>& rows,
                        Buf<int>& levptr, int& nlev) {
    lev.alloc(c, n); lev.zero();
    Buf<int> flag(c, 2);
    for (int sweep = 0; sweep <= n; ++sweep) {
        flag.zero();
        SSN_LAUNCH(c, level_relax_kernel, cdiv(n, 256), 256, 0, n, p, idx, diag_first, lev.p, flag.p);
        if (read_scalar(c, flag.p) == 0) break;
    }
    Buf<int> cnt(c, (size_t)n + 1); cnt.zero(); flag.zero();
    SSN_LAUNCH(c, level_hist_kernel, cdiv(n, 256), 256, 0, n, lev.p, cnt.p, flag.p);
    nlev = read_scalar(c, flag.p) + 1;
    levptr.alloc(c, (size_t)nlev + 1);
    scan_counts_to_ptr(c, cnt.p, levptr.p, nlev);
    Buf<int> ids(c, n), keys_out(c, n);
    rows.alloc(c, n);
    iota_int(c, ids.p, n);
    stable_sort_pairs(c, lev.p, keys_out.p, ids.p, rows.p, n, nlev > 1 ? nlev : 2);
}

}  // namespace

void build_tri_factors_device(ssn_ctx* c, const CsrView& H, int precd, TriFactors& F) {
    const int n = H.nrows;
    SSN_REQUIRE(n > 0, SSN_E_INVALID, "PCG: empty matrix");
    Buf<int> lcnt(c, n), ucnt(c, n), flag(c, 1);
    flag.zero();
    SSN_LAUNCH(c, tri_count_kernel, cdiv(n, 256), 256, 0, n, H.ptr, H.idx, lcnt.p, ucnt.p, flag.p);
    F.lp.alloc(c, (size_t)n + 1);
    const int64_t nnzl = scan_counts_to_ptr(c, lcnt.p, F.lp.p, n);
    F.li.alloc(c, nnzl); F.lv.alloc(c, nnzl);
    F.has_mid = false;
    if (precd == 3) {
        F.up.alloc(c, (size_t)n + 1);
        const int64_t nnzu = scan_counts_to_ptr(c, ucnt.p, F.up.p, n);
        F.ui.alloc(c, nnzu); F.uv.alloc(c, nnzu); F.mid.alloc(c, n); F.has_mid = true;
        const double om = 1.5, sc = om * (2.0 - om);
        SSN_LAUNCH(c, ssor_fill_kernel, cdiv(n, 256), 256, 0, n, H.ptr, H.idx, H.val, om, sc, F.lp.p, F.li.p, F.lv.p,
                   F.up.p, F.ui.p, F.uv.p, F.mid.p);
        Buf<int> lev;
        levels_and_buckets(c, n, F.lp.p, F.li.p, 0, lev, F.lrows, F.llev, F.nl);
    } else {
        SSN_REQUIRE(read_scalar(c, flag.p) == 0, SSN_E_NOT_SPD, "ichol: zero on the diagonal");
        SSN_LAUNCH(c, tril_fill_kernel, cdiv(n, 256), 256, 0, n, H.ptr, H.idx, H.val, F.lp.p, F.li.p, F.lv.p);
        Buf<int> lev;
        levels_and_buckets(c, n, F.lp.p, F.li.p, 0, lev, F.lrows, F.llev, F.nl);
        std::vector<int> hlev((size_t)F.nl + 1);
        read_back(c, F.llev.p, hlev.data(), hlev.size());
        flag.zero();
        for (int l = 0; l < F.nl; ++l) {
            const int cntl = hlev[(size_t)l + 1] - hlev[l];
            if (cntl > 0) SSN_LAUNCH(c, ic0_level_kernel, cdiv(cntl, 128), 128, 0, cntl, F.lrows.p + hlev[l], F.lp.p, F.li.p, F.lv.p, flag.p);
        }
        SSN_REQUIRE(read_scalar(c, flag.p) == 0, SSN_E_NOT_SPD, "ichol: encountered nonpositive pivot");
        // Uf = Lf' (stable sort by column: rows ascending inside a column, so the diagonal comes first)
        CsrView Lv; Lv.nrows = n; Lv.ncols = n; Lv.nnz = nnzl; Lv.ptr = F.lp.p; Lv.idx = F.li.p; Lv.val = F.lv.p;
        Csr U = transpose(c, Lv);
        F.up = std::move(U.ptr); F.ui = std::move(U.idx); F.uv = std::move(U.val);
    }
    Buf<int> lev;
    levels_and_buckets(c, n, F.up.p, F.ui.p, 1, lev, F.urows, F.ulev, F.nu);
}

// Jk = bk1*speye(m+n) + (T + H0)/tk -- Class1/APD_SsN_Class1.m:147,151 (the matrix PCG is given for inner_solver = 2);
// `/tk` is a product with 1/tk like everywhere on the path (the oracle's frozen convention)
namespace {

__global__ void jk_fill_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                               const double* __restrict__ t, double bk1, double inv_tk, const int* __restrict__ optr,
                               int* __restrict__ oidx, double* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    const int e0 = ptr[row], e1 = ptr[row + 1], o = optr[row];
    const bool has = (optr[row + 1] - o) == (e1 - e0);
    const double tv = t ? t[row] : 0.0;
    int nlow = 0;
    for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        int j = 0x7fffffff; double h = 0.0;
        if (e < e1) { j = idx[e]; h = val[e]; }
        const unsigned lowmask = __ballot_sync(0xffffffffu, j < row);
        if (e < e1) {
            const int pos = o + (e - e0) + ((!has && j > row) ? 1 : 0);
            oidx[pos] = j;
            oval[pos] = (j == row) ? __dadd_rn(bk1, __dmul_rn(inv_tk, __dadd_rn(tv, h))) : __dmul_rn(inv_tk, h);
        }
        nlow += __popc(lowmask);
    }
    if (!has && lane == 0) {
        oidx[o + nlow] = row;
        oval[o + nlow] = __dadd_rn(bk1, __dmul_rn(inv_tk, tv));
    }
}

__global__ void jk_count_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ idx, int* __restrict__ len) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    bool has = false;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) has |= (idx[e] == row);
    has = __any_sync(0xffffffffu, has);
    if (lane == 0) len[row] = ptr[row + 1] - ptr[row] + (has ? 0 : 1);
}

}  // namespace

void jk_system(ssn_ctx* c, const ssn_prob_data* pd, Csr& Jk) {
    SSN_REQUIRE(pd && pd->H0, SSN_E_INVALID, "prob_data: missing field");
    const int N = (int)(pd->n + pd->m);
    CsrView H0(*pd->H0);
    SSN_REQUIRE(H0.nrows == N && H0.ncols == N, SSN_E_INVALID, "prob_data.H0 must be (n+m) x (n+m)");
    Buf<int> len(c, N);
    SSN_LAUNCH(c, jk_count_kernel, cdiv((int64_t)N * 32, 256), 256, 0, N, H0.ptr, H0.idx, len.p);
    Jk = csr_alloc_from_counts(c, N, N, len);
    SSN_LAUNCH(c, jk_fill_kernel, cdiv((int64_t)N * 32, 256), 256, 0, N, H0.ptr, H0.idx, H0.val, pd->t_dev, pd->bk1, 1.0 / pd->tk,
               Jk.ptr.p, Jk.idx.p, Jk.val.p);
}

}  // namespace ssn
