// solvers.cu -- ASAt assembly (K4 back half), the rescaled system (K5), connected components
// (K6), the batched small-component direct solve (K16), Hybrid_AMG, aug_PCG, the POT bordering
// and the closed-form inverses invAAt / invHHt.
#include "amg.cuh"
#include "plan_ops.cuh"
#include "solvers.cuh"

#include <algorithm>

namespace ssn {

// =================================================================== ASAt

namespace {

// diag value and row length of H for one node; thread per node, sequential ascending sums that
// start from 0.0 (the order of MATLAB's U'*p and Q*q mat-vecs, ASAt.m:19)
__global__ void asat_rowlen_kernel(int n, int m, const int* __restrict__ colptr, const int* __restrict__ yrow,
                                   const int* __restrict__ rowptr, const int* __restrict__ ycolT,
                                   const double* __restrict__ p, const double* __restrict__ q,
                                   double* __restrict__ dval, int* __restrict__ len) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n + m) return;
    double s = 0.0; int cnt;
    if (v < n) {
        const int e0 = colptr[v], e1 = colptr[v + 1];
        for (int e = e0; e < e1; ++e) { const double pi = p[yrow[e]]; s = __dadd_rn(s, __dmul_rn(pi, pi)); }
        cnt = e1 - e0;
    } else {
        const int i = v - n;
        const int e0 = rowptr[i], e1 = rowptr[i + 1];
        for (int e = e0; e < e1; ++e) { const double qj = q[ycolT[e]]; s = __dadd_rn(s, __dmul_rn(qj, qj)); }
        cnt = e1 - e0;
    }
    dval[v] = s;
    len[v] = cnt + (s != 0.0 ? 1 : 0);
}

__global__ void asat_fill_kernel(int n, int m, const int* __restrict__ colptr, const int* __restrict__ yrow,
                                 const int* __restrict__ rowptr, const int* __restrict__ ycolT,
                                 const double* __restrict__ p, const double* __restrict__ q,
                                 const double* __restrict__ dval, const int* __restrict__ hptr,
                                 int* __restrict__ hidx, double* __restrict__ hval) {
    const int lane = threadIdx.x & 31;
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (v >= n + m) return;
    int o = hptr[v];
    const double d = dval[v];
    if (v < n) {                                         // [diag , q_j * p_i at column n+i]
        if (d != 0.0) { if (lane == 0) { hidx[o] = v; hval[o] = d; } ++o; }
        const int e0 = colptr[v], len = colptr[v + 1] - e0;
        const double qj = q[v];
        for (int t = lane; t < len; t += 32) { const int i = yrow[e0 + t]; hidx[o + t] = n + i; hval[o + t] = __dmul_rn(qj, p[i]); }
    } else {                                             // [p_i * q_j at column j , diag]
        const int i = v - n;
        const int e0 = rowptr[i], len = rowptr[i + 1] - e0;
        const double pi = p[i];
        for (int t = lane; t < len; t += 32) { const int j = ycolT[e0 + t]; hidx[o + t] = j; hval[o + t] = __dmul_rn(pi, q[j]); }
        if (d != 0.0 && lane == 0) { hidx[o + len] = v; hval[o + len] = d; }
    }
}

// ASAtz.m:15-22 from the compacted active set; thread per node, sequential sums
__global__ void asatz_kernel(int n, int m, const int* __restrict__ colptr, const int* __restrict__ yrow,
                             const int* __restrict__ rowptr, const int* __restrict__ ycolT,
                             const double* __restrict__ p, const double* __restrict__ q,
                             const double* __restrict__ z, double* __restrict__ y) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n + m) return;
    if (v < n) {                                         // y1 = (U'p).*z1 + q.*(U'z2)
        double upp = 0.0, uz = 0.0;
        for (int e = colptr[v]; e < colptr[v + 1]; ++e) { const int i = yrow[e]; upp = fma(p[i], p[i], upp); uz = fma(p[i], z[n + i], uz); }
        y[v] = upp * z[v] + q[v] * uz;
    } else {                                             // y2 = p.*(Q z1) + (Q p).*z2   (Q*p as written, ASAtz.m:21)
        const int i = v - n;
        double qz = 0.0, qp = 0.0;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) { const int j = ycolT[e]; qz = fma(q[j], z[j], qz); qp = fma(q[j], p[j], qp); }
        y[v] = p[i] * qz + qp * z[v];
    }
}

struct ActiveSet { Buf<int> colptr, yrow, ycol, rowptr, ycolT; int64_t E = 0; };

void build_active_set(ssn_ctx* c, const uint8_t* s, int64_t m, int64_t n, ActiveSet& a) {
    Buf<int> rowcount;
    a.E = plan_active_set(c, s, m, n, a.colptr, a.yrow, a.ycol, rowcount);
    a.rowptr.alloc(c, m + 1);
    scan_counts_async(c, rowcount, a.rowptr, m);                     // the total is a.E: no host read
    a.ycolT.alloc(c, a.E);
    if (a.E > 0) {
        Buf<int> keys_out(c, a.E);
        stable_sort_pairs(c, a.yrow, keys_out, a.ycol, a.ycolT, a.E, (int)(m > 1 ? m : 2));   // CSR of Y: ascending j per row
    }
}

}  // namespace

namespace {

__global__ void lin_from_coo_kernel(int64_t E, const int* __restrict__ yrow, const int* __restrict__ ycol, int64_t row_offset,
                                    int64_t m_global, long long* __restrict__ lin) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
        lin[e] = (long long)(yrow[e] + row_offset) + (long long)ycol[e] * (long long)m_global;
}
__global__ void coo_from_lin_kernel(int64_t E, const long long* __restrict__ lin, int64_t m, int* __restrict__ yrow,
                                    int* __restrict__ ycol, int* __restrict__ colcount, int* __restrict__ rowcount) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const long long l = lin[e];
        const int j = (int)(l / m), i = (int)(l - (long long)j * m);
        yrow[e] = i; ycol[e] = j;
        atomicAdd(colcount + j, 1); atomicAdd(rowcount + i, 1);
    }
}

Csr asat_from_active(ssn_ctx* c, ActiveSet& a, const double* p, const double* q, int64_t m, int64_t n) {
    const int N = (int)(m + n);
    Buf<double> dval(c, N); Buf<int> len(c, N);
    SSN_LAUNCH(c, asat_rowlen_kernel, cdiv(N, 128), 128, 0, (int)n, (int)m, a.colptr.p, a.yrow.p, a.rowptr.p, a.ycolT.p, p, q, dval.p, len.p);
    Csr H = csr_alloc_from_counts(c, N, N, len);
    if (H.nnz) SSN_LAUNCH(c, asat_fill_kernel, cdiv((int64_t)N * 32, 256), 256, 0, (int)n, (int)m, a.colptr.p, a.yrow.p, a.rowptr.p,
                          a.ycolT.p, p, q, dval.p, H.ptr.p, H.idx.p, H.val.p);
    return H;
}

}  // namespace

// The active set of a (row slab of a) plan as global column-major linear indices, in the slab's
// own CSC order (used by the row-sharded multi-GPU path to exchange active sets, O(E) integers).
void active_coo(ssn_ctx* c, const uint8_t* s, int64_t m_loc, int64_t n, int64_t row_offset, int64_t m_global,
                long long** lin_out, int64_t* E_out) {
    Buf<int> colptr, yrow, ycol, rowcount;
    const int64_t E = plan_active_set(c, s, m_loc, n, colptr, yrow, ycol, rowcount);
    Buf<long long> lin(c, E);
    if (E) SSN_LAUNCH(c, lin_from_coo_kernel, 592, 256, 0, E, yrow.p, ycol.p, row_offset, m_global, lin.p);
    *E_out = E; *lin_out = lin.release();
}

// H = ASAt from the sorted global linear indices of the active set (what find(s) returns)
Csr asat_coo(ssn_ctx* c, const long long* lin_sorted, int64_t E, const double* p, const double* q, int64_t m, int64_t n) {
    SSN_REQUIRE(p && q && m > 0 && n > 0 && E >= 0 && (E == 0 || lin_sorted), SSN_E_INVALID, "ASAt(coo): bad arguments");
    SSN_REQUIRE(E < ((int64_t)1 << 30), SSN_E_TOO_LARGE, "ASAt(coo): nnz(s) >= 2^30");
    ActiveSet a; a.E = E;
    Buf<int> colcount(c, n), rowcount(c, m);
    colcount.zero(); rowcount.zero();
    a.yrow.alloc(c, E); a.ycol.alloc(c, E); a.colptr.alloc(c, n + 1); a.rowptr.alloc(c, m + 1); a.ycolT.alloc(c, E);
    if (E) SSN_LAUNCH(c, coo_from_lin_kernel, 592, 256, 0, E, lin_sorted, m, a.yrow.p, a.ycol.p, colcount.p, rowcount.p);
    scan_counts_async(c, colcount, a.colptr, n);                     // both totals are E: no host reads
    scan_counts_async(c, rowcount, a.rowptr, m);
    if (E) {
        Buf<int> keys_out(c, E);
        stable_sort_pairs(c, a.yrow, keys_out, a.ycol, a.ycolT, E, (int)(m > 1 ? m : 2));
    }
    return asat_from_active(c, a, p, q, m, n);
}

Csr asat(ssn_ctx* c, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n) {
    SSN_REQUIRE(s && p && q && m > 0 && n > 0, SSN_E_INVALID, "ASAt: bad arguments");
    SSN_REQUIRE(m + n < ((int64_t)1 << 30), SSN_E_TOO_LARGE, "ASAt: m+n too large");
    ActiveSet a;
    build_active_set(c, s, m, n, a);
    const int N = (int)(m + n);
    Buf<double> dval(c, N); Buf<int> len(c, N);
    SSN_LAUNCH(c, asat_rowlen_kernel, cdiv(N, 128), 128, 0, (int)n, (int)m, a.colptr.p, a.yrow.p, a.rowptr.p, a.ycolT.p, p, q, dval.p, len.p);
    Csr H = csr_alloc_from_counts(c, N, N, len);
    if (H.nnz) SSN_LAUNCH(c, asat_fill_kernel, cdiv((int64_t)N * 32, 256), 256, 0, (int)n, (int)m, a.colptr.p, a.yrow.p, a.rowptr.p,
                          a.ycolT.p, p, q, dval.p, H.ptr.p, H.idx.p, H.val.p);
    return H;
}

void asatz(ssn_ctx* c, const double* z, const uint8_t* s, const double* p, const double* q, int64_t m, int64_t n, double* y) {
    SSN_REQUIRE(z && s && p && q && y && m > 0 && n > 0, SSN_E_INVALID, "ASAtz: bad arguments");
    SSN_REQUIRE(m == n, SSN_E_ASATZ_DIM, "ASAtz.m:21 multiplies the m-by-n matrix Q by p (length m): needs m == n");
    ActiveSet a;
    build_active_set(c, s, m, n, a);
    const int N = (int)(m + n);
    SSN_LAUNCH(c, asatz_kernel, cdiv(N, 128), 128, 0, (int)n, (int)m, a.colptr.p, a.yrow.p, a.rowptr.p, a.ycolT.p, p, q, z, y);
}

// =================================================================== rescaled system (K5)

namespace {

__device__ __forceinline__ double qp_of(int v, int n, const double* __restrict__ p, const double* __restrict__ q) {
    return v < n ? q[v] : -p[v - n];
}

__global__ void qp_kernel(int n, int m, const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ t,
                          const double* __restrict__ z, double* __restrict__ qp, double* __restrict__ Kd, double* __restrict__ f,
                          int* __restrict__ zero_flag) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n + m) return;
    const double s = qp_of(v, n, p, q);
    if (s == 0.0) *zero_flag = 1;                                  // Hybrid_AMG.m:18-19
    qp[v] = s;
    const double tv = t ? t[v] : 0.0;
    Kd[v] = __dmul_rn(__dmul_rn(s, tv), s);                        // K = Q0*T*Q0
    if (f) f[v] = __dmul_rn(s, z[v]);                              // f = Q0*z
}

__global__ void ae_count_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ idx, int* __restrict__ len) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    bool has = false;
    for (int e = ptr[row] + lane; e < ptr[row + 1]; e += 32) has |= (idx[e] == row);
    has = __any_sync(0xffffffffu, has);
    if (lane == 0) len[row] = ptr[row + 1] - ptr[row] + (has ? 0 : 1);
}

// Ae = bk1*Q + (1/tk)*(K + A0), A0_ij = (qp_i*h_ij)*qp_j; the diagonal is always stored.
__global__ void ae_fill_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                               const double* __restrict__ qp, const double* __restrict__ Kd, double bk1, double inv_tk,
                               const int* __restrict__ optr, int* __restrict__ oidx, double* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    const int e0 = ptr[row], e1 = ptr[row + 1], o = optr[row];
    const bool has = (optr[row + 1] - o) == (e1 - e0);
    const double si = qp[row];
    const double qd = __dmul_rn(bk1, __dmul_rn(si, si));
    int nlow = 0;
    for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        int j = 0x7fffffff; double h = 0.0;
        if (e < e1) { j = idx[e]; h = val[e]; }
        const unsigned lowmask = __ballot_sync(0xffffffffu, j < row);
        if (e < e1) {
            const double a0 = __dmul_rn(__dmul_rn(si, h), qp[j]);
            const int pos = o + (e - e0) + ((!has && j > row) ? 1 : 0);
            oidx[pos] = j;
            oval[pos] = (j == row) ? __dadd_rn(qd, __dmul_rn(inv_tk, __dadd_rn(Kd[row], a0))) : __dmul_rn(inv_tk, a0);
        }
        nlow += __popc(lowmask);
    }
    if (!has && lane == 0) {
        oidx[o + nlow] = row;
        oval[o + nlow] = __dadd_rn(qd, __dmul_rn(inv_tk, Kd[row]));
    }
}

}  // namespace

void rescaled_system(ssn_ctx* c, const ssn_prob_data* pd, Csr& Ae, double* f, Buf<double>& qp, Buf<double>& Kd) {
    SSN_REQUIRE(pd && pd->H0 && pd->p_dev && pd->q_dev, SSN_E_INVALID, "prob_data: missing field");
    const int n = (int)pd->n, m = (int)pd->m, N = n + m;
    CsrView H0(*pd->H0);
    SSN_REQUIRE(H0.nrows == N && H0.ncols == N, SSN_E_INVALID, "prob_data.H0 must be (n+m) x (n+m)");
    qp.alloc(c, N); Kd.alloc(c, N);
    Buf<int> zflag(c, 1); zflag.zero();
    SSN_LAUNCH(c, qp_kernel, cdiv(N, 256), 256, 0, n, m, pd->p_dev, pd->q_dev, pd->t_dev, pd->z_dev, qp.p, Kd.p,
               (pd->z_dev ? f : nullptr), zflag.p);
    SSN_REQUIRE(read_scalar(c, zflag.p) == 0, SSN_E_PQ_ZERO, "p or q contains 0 !!!!!");
    Buf<int> len(c, N);
    SSN_LAUNCH(c, ae_count_kernel, cdiv((int64_t)N * 32, 256), 256, 0, N, H0.ptr, H0.idx, len.p);
    Ae = csr_alloc_from_counts(c, N, N, len);
    SSN_LAUNCH(c, ae_fill_kernel, cdiv((int64_t)N * 32, 256), 256, 0, N, H0.ptr, H0.idx, H0.val, qp.p, Kd.p, pd->bk1,
               1.0 / pd->tk, Ae.ptr.p, Ae.idx.p, Ae.val.p);
}

// =================================================================== components (K6)

namespace {

__global__ void cc_hook_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                               int* __restrict__ comp, int* __restrict__ changed) {
    const int lane = threadIdx.x & 31;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= n) return;
    for (int e = ptr[u] + lane; e < ptr[u + 1]; e += 32) {
        const int v = idx[e];
        if (v == u || val[e] == 0.0) continue;
        const int cu = comp[u], cv = comp[v];
        if (cu < cv) { atomicMin(comp + cv, cu); *changed = 1; }
        else if (cv < cu) { atomicMin(comp + cu, cv); *changed = 1; }
    }
}
__global__ void cc_jump_kernel(int n, int* __restrict__ comp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ci = comp[i];
    while (true) { const int cc = comp[ci]; if (cc == ci) break; ci = cc; }
    comp[i] = ci;
}
__global__ void cc_root_kernel(int n, const int* __restrict__ comp, int* __restrict__ isroot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) isroot[i] = comp[i] == i ? 1 : 0;
}
__global__ void cc_label_kernel(int n, const int* __restrict__ comp, const int* __restrict__ rootrank,
                                int* __restrict__ blocks0, int* __restrict__ blocks1, int* __restrict__ sizes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = rootrank[comp[i]];
    blocks0[i] = b; if (blocks1) blocks1[i] = b + 1;
    atomicAdd(sizes + b, 1);
}

}  // namespace

// blocks: 1-based labels; sizes[ncomp]; perm: nodes grouped by component (ascending inside);
// r: block boundaries (ncomp+1).  Component order = ascending smallest member.
void components(ssn_ctx* c, const CsrView& A, int* blocks, int* sizes, int* perm, int* r, int* ncomp_out) {
    SSN_REQUIRE(A.nrows == A.ncols, SSN_E_NOT_SQUARE, "Adjacency matrix must be square");
    const int n = A.nrows;
    if (n == 0) { if (ncomp_out) *ncomp_out = 0; return; }
    Buf<int> comp(c, n), changed(c, 1);
    iota_int(c, comp, n);
    const int g = cdiv(n, 256), gw = cdiv((int64_t)n * 32, 256);
    for (int iter = 0; iter < 10000; ++iter) {
        changed.zero();
        SSN_LAUNCH(c, cc_hook_kernel, gw, 256, 0, n, A.ptr, A.idx, A.val, comp.p, changed.p);
        SSN_LAUNCH(c, cc_jump_kernel, g, 256, 0, n, comp.p);
        if (read_scalar(c, changed.p) == 0) break;
    }
    Buf<int> isroot(c, n), rootrank(c, (size_t)n + 1), blocks0(c, n), keys_out(c, n), ids(c, n);
    SSN_LAUNCH(c, cc_root_kernel, g, 256, 0, n, comp.p, isroot.p);
    const int ncomp = (int)scan_counts_to_ptr(c, isroot, rootrank, n);
    SSN_CUDA(cudaMemsetAsync(sizes, 0, sizeof(int) * n, c->stream));
    SSN_LAUNCH(c, cc_label_kernel, g, 256, 0, n, comp.p, rootrank.p, blocks0.p, blocks, sizes);
    iota_int(c, ids, n);
    stable_sort_pairs(c, blocks0, keys_out, ids, perm, n, ncomp > 1 ? ncomp : 2);
    scan_counts_async(c, sizes, r, ncomp);                           // the total is n: no host read
    if (ncomp_out) *ncomp_out = ncomp;
}

// =================================================================== small components (K16)

namespace {

__global__ void scatter_pos_kernel(int n, const int* __restrict__ perm, const int* __restrict__ blocks0_of_perm_unused,
                                   int* __restrict__ pos) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) pos[perm[t]] = t;
}

// one block per listed component: dense Cholesky in shared memory, then two triangular solves
__global__ void __launch_bounds__(128) small_chol_kernel(const int* __restrict__ comp_list, const int* __restrict__ r,
                                                         const int* __restrict__ perm, const int* __restrict__ pos,
                                                         const int* __restrict__ ptr, const int* __restrict__ idx,
                                                         const double* __restrict__ val, const double* __restrict__ f,
                                                         double* __restrict__ u, int* __restrict__ err) {
    extern __shared__ double sm[];
    const int cid = comp_list[blockIdx.x];
    const int r0 = r[cid], sz = r[cid + 1] - r0;
    double* Am = sm;                    // sz x sz, row-major
    double* b = sm + sz * sz;           // sz
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int t = tid; t < sz * sz; t += nt) Am[t] = 0.0;
    __syncthreads();
    for (int li = tid; li < sz; li += nt) {
        const int v = perm[r0 + li];
        b[li] = f[v];
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) Am[li * sz + (pos[idx[e]] - r0)] = val[e];
    }
    __syncthreads();
    for (int k = 0; k < sz; ++k) {
        if (tid == 0) {
            const double d = Am[k * sz + k];
            if (!(d > 0.0)) *err = 1;
            Am[k * sz + k] = sqrt(d);
        }
        __syncthreads();
        const double lkk = Am[k * sz + k];
        for (int i = k + 1 + tid; i < sz; i += nt) Am[i * sz + k] /= lkk;
        __syncthreads();
        const int rem = sz - k - 1;
        for (int t = tid; t < rem * rem; t += nt) {
            const int i = k + 1 + t / rem, j = k + 1 + t % rem;
            if (j <= i) Am[i * sz + j] -= Am[i * sz + k] * Am[j * sz + k];
        }
        __syncthreads();
    }
    // L y = b ; L' x = y  (serial in k, parallel over rows)
    for (int k = 0; k < sz; ++k) {
        if (tid == 0) b[k] /= Am[k * sz + k];
        __syncthreads();
        const double yk = b[k];
        for (int i = k + 1 + tid; i < sz; i += nt) b[i] -= Am[i * sz + k] * yk;
        __syncthreads();
    }
    for (int k = sz - 1; k >= 0; --k) {
        if (tid == 0) b[k] /= Am[k * sz + k];
        __syncthreads();
        const double xk = b[k];
        for (int i = tid; i < k; i += nt) b[i] -= Am[k * sz + i] * xk;
        __syncthreads();
    }
    for (int li = tid; li < sz; li += nt) u[perm[r0 + li]] = b[li];
}

__global__ void singleton_solve_kernel(int n, const int* __restrict__ blocks1, const int* __restrict__ sizes,
                                       const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                                       const double* __restrict__ f, double* __restrict__ u) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n || sizes[blocks1[v] - 1] != 1) return;
    double d = 0.0;
    for (int e = ptr[v]; e < ptr[v + 1]; ++e) if (idx[e] == v) d = val[e];
    u[v] = f[v] / d;
}

__global__ void gather_kernel(int n, const int* __restrict__ sel, const double* __restrict__ x, double* __restrict__ y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) y[t] = x[sel[t]];
}
__global__ void scatter_kernel(int n, const int* __restrict__ sel, const double* __restrict__ x, double* __restrict__ y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) y[sel[t]] = x[t];
}
__global__ void scatter_index_kernel(int n, const int* __restrict__ sel, int* __restrict__ newidx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) newidx[sel[t]] = t;
}
__global__ void scale_kernel(int n, double a, double* __restrict__ x) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t] = __dmul_rn(a, x[t]);
}
__global__ void mul_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) y[t] = a[t] * b[t];
}
__global__ void count_less_kernel(int n, const int* __restrict__ sel, int bound, int* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int v = (t < n && sel[t] < bound) ? 1 : 0;
    v = warp_sum_int(v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

}  // namespace

// =================================================================== Hybrid_AMG

// Hybrid_AMG.m and Hybrid_twogrid.m are the same dispatch (rescaled system, components, large components by the
// multilevel solver with a random guess, small ones directly); they differ in the solver: Class_AMG
// (Hybrid_AMG.m:41,70) or twogrid_bigph (Hybrid_twogrid.m:39,67).
void hybrid_amg(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, double* zeta, int* itamg_out,
                double* resamg_out, int* info, bool twogrid) {
    SSN_REQUIRE(pd && zeta && pd->z_dev, SSN_E_INVALID, "Hybrid_AMG: bad arguments");
    const int n = (int)pd->n, m = (int)pd->m, N = n + m;
    const double bk1 = pd->bk1, tk = pd->tk;
    Csr Ae; Buf<double> f(c, N), qp, Kd, u(c, N);
    { Phase ph(c, "hybrid.rescaled_system"); rescaled_system(c, pd, Ae, f, qp, Kd); }                                   // Hybrid_AMG.m:17-24
    Buf<int> blocks(c, N), sizes(c, N), perm(c, N), r(c, (size_t)N + 1);
    int ncomp = 0;
    { Phase ph(c, "hybrid.components"); components(c, Ae, blocks, sizes, perm, r, &ncomp); }                       // :27
    AmgOptions o = resolve_options(opts);
    int itamg = 0, it_num = 0; double resamg = 0.0;
    const double gscale = bk1 * tk;
    if (ncomp == 1) {                                                         // :30-48
        o.isnsp = (dev_sum(c, Kd, N) != 0.0) ? 0 : 1;
        o.fnode = n;
        Buf<double> guess(c, N);
        rng_rand(c, N, guess);                                                // :40
        SSN_LAUNCH(c, scale_kernel, cdiv(N, 256), 256, 0, N, gscale, guess.p);
        o.guess = guess;
        if (twogrid) twogrid_bigph(c, Ae, f, o, u, &itamg, &resamg, nullptr, nullptr, nullptr);
        else         class_amg(c, Ae, f, o, false, u, &itamg, &resamg, nullptr, nullptr, nullptr);
        it_num = 1;
    } else {                                                                  // :50-107
        std::vector<int> hs(ncomp), hr(ncomp + 1);
        read_back(c, sizes.p, hs.data(), (size_t)ncomp);
        read_back(c, r.p, hr.data(), (size_t)ncomp + 1);
        fill_double(c, u, N, 0.0);
        const int N0 = 100;
        Buf<int> newidx(c, N);
        std::vector<int> small_multi;
        bool any_single = false;
        for (int k = 0; k < ncomp; ++k) {
            const int sz = hs[k];
            if (sz <= N0) { if (sz == 1) any_single = true; else small_multi.push_back(k); continue; }
            const int* pk = perm.p + hr[k];
            fill_int(c, newidx, N, -1);
            SSN_LAUNCH(c, scatter_index_kernel, cdiv(sz, 256), 256, 0, sz, pk, newidx.p);
            Csr Aek = extract_principal(c, Ae, pk, sz, newidx);              // Ae(pk,pk)
            Buf<double> fk(c, sz), dKk(c, sz), dk(c, sz), guess(c, sz);
            SSN_LAUNCH(c, gather_kernel, cdiv(sz, 256), 256, 0, sz, pk, f.p, fk.p);
            SSN_LAUNCH(c, gather_kernel, cdiv(sz, 256), 256, 0, sz, pk, Kd.p, dKk.p);
            o.isnsp = (dev_sum(c, dKk, sz) != 0.0) ? 0 : 1;                   // :60-66
            Buf<int> cnt(c, 1); cnt.zero();
            SSN_LAUNCH(c, count_less_kernel, cdiv(sz, 256), 256, 0, sz, pk, n, cnt.p);
            o.fnode = read_scalar(c, cnt.p);                                  // :68
            rng_rand(c, sz, guess);                                           // :69
            SSN_LAUNCH(c, scale_kernel, cdiv(sz, 256), 256, 0, sz, gscale, guess.p);
            o.guess = guess;
            int itk = 0; double resk = 0.0;
            if (twogrid) twogrid_bigph(c, Aek, fk, o, dk, &itk, &resk, nullptr, nullptr, nullptr);
            else         class_amg(c, Aek, fk, o, false, dk, &itk, &resk, nullptr, nullptr, nullptr);
            SSN_LAUNCH(c, scatter_kernel, cdiv(sz, 256), 256, 0, sz, pk, dk.p, u.p);
            itamg = std::max(itamg, itk); resamg = std::max(resamg, resk);
            it_num = k + 1;                                                   // :80
        }
        if (any_single)                                                       // :85-91, 1x1 blocks
            SSN_LAUNCH(c, singleton_solve_kernel, cdiv(N, 256), 256, 0, N, blocks.p, sizes.p, Ae.ptr.p, Ae.idx.p, Ae.val.p, f.p, u.p);
        if (!small_multi.empty()) {
            Buf<int> list(c, small_multi.size()), pos(c, N), err(c, 1);
            err.zero();
            upload_small(c, list.p, small_multi.data(), sizeof(int) * small_multi.size());
            SSN_LAUNCH(c, scatter_pos_kernel, cdiv(N, 256), 256, 0, N, perm.p, nullptr, pos.p);
            int maxsz = 2;
            for (int k : small_multi) maxsz = std::max(maxsz, hs[k]);
            const size_t smem = sizeof(double) * ((size_t)maxsz * maxsz + maxsz);
            SSN_CUDA(cudaFuncSetAttribute(small_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SSN_LAUNCH(c, small_chol_kernel, (int)small_multi.size(), 128, smem, list.p, r.p, perm.p, pos.p, Ae.ptr.p, Ae.idx.p,
                       Ae.val.p, f.p, u.p, err.p);
            SSN_CUDA(cudaStreamSynchronize(c->stream));      // list's host source must outlive the copy
            SSN_REQUIRE(read_scalar(c, err.p) == 0, SSN_E_NOT_SPD, "small-component Cholesky: non-positive pivot");
        }
    }
    SSN_LAUNCH(c, mul_kernel, cdiv(N, 256), 256, 0, N, qp.p, u.p, zeta);     // zeta = Q0*u, :113
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    if (itamg_out) *itamg_out = itamg;
    if (resamg_out) *resamg_out = resamg;
    if (info) { info[0] = ncomp; info[1] = it_num; }
}

// =================================================================== aug_PCG

namespace {

__global__ void qk_kernel(int N, const double* __restrict__ qp, const double* __restrict__ Kd, double bk1, double inv_tk,
                          double* __restrict__ QK) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) QK[v] = bk1 * (qp[v] * qp[v]) + inv_tk * Kd[v];              // aug_PCG.m:27
}
// rows [0,nc): [ (Y'QKY)_cc , QK_i at nc+i for members i ] ; rows nc+i: [ QK_i at blocks_i , Ae(i,:) shifted ]
__global__ void aug_count_kernel(int nc, int N, const int* __restrict__ sizes, const int* __restrict__ aptr, int* __restrict__ len) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nc) len[t] = 1 + sizes[t];
    else if (t < nc + N) len[t] = 1 + aptr[t - nc + 1] - aptr[t - nc];
}
__global__ void aug_fill_kernel(int nc, int N, const int* __restrict__ r, const int* __restrict__ perm,
                                const int* __restrict__ blocks1, const double* __restrict__ QK, const double* __restrict__ f,
                                const int* __restrict__ aptr, const int* __restrict__ aidx, const double* __restrict__ aval,
                                const int* __restrict__ optr, int* __restrict__ oidx, double* __restrict__ oval,
                                double* __restrict__ augf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nc) {
        int o = optr[t];
        double s = 0.0, sf = 0.0;
        for (int e = r[t]; e < r[t + 1]; ++e) { const int i = perm[e]; s += QK[i]; sf += f[i]; }
        oidx[o] = t; oval[o] = s; ++o;
        for (int e = r[t]; e < r[t + 1]; ++e) { const int i = perm[e]; oidx[o] = nc + i; oval[o] = QK[i]; ++o; }
        augf[t] = sf;
    } else if (t < nc + N) {
        const int i = t - nc;
        int o = optr[t];
        oidx[o] = blocks1[i] - 1; oval[o] = QK[i]; ++o;
        for (int e = aptr[i]; e < aptr[i + 1]; ++e) { oidx[o] = nc + aidx[e]; oval[o] = aval[e]; ++o; }
        augf[t] = f[i];
    }
}
__global__ void aug_recombine_kernel(int nc, int N, const int* __restrict__ blocks1, const double* __restrict__ U,
                                     const double* __restrict__ qp, double* __restrict__ zeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) zeta[i] = qp[i] * (U[blocks1[i] - 1] + U[nc + i]);            // aug_PCG.m:35-37
}

}  // namespace

void aug_pcg(ssn_ctx* c, const ssn_prob_data* pd, const ssn_pcg_options* opts, double* zeta, int* itpcg, double* respcg, int* info) {
    SSN_REQUIRE(pd && zeta && pd->z_dev, SSN_E_INVALID, "aug_PCG: bad arguments");
    const int n = (int)pd->n, m = (int)pd->m, N = n + m;
    Csr Ae; Buf<double> f(c, N), qp, Kd;
    rescaled_system(c, pd, Ae, f, qp, Kd);
    Buf<int> blocks(c, N), sizes(c, N), perm(c, N), r(c, (size_t)N + 1);
    int nc = 0;
    components(c, Ae, blocks, sizes, perm, r, &nc);                           // aug_PCG.m:24
    Buf<double> QK(c, N);
    SSN_LAUNCH(c, qk_kernel, cdiv(N, 256), 256, 0, N, qp.p, Kd.p, pd->bk1, 1.0 / pd->tk, QK.p);
    const int M = nc + N;
    Buf<int> len(c, M);
    SSN_LAUNCH(c, aug_count_kernel, cdiv(M, 256), 256, 0, nc, N, sizes.p, Ae.ptr.p, len.p);
    Csr Aug = csr_alloc_from_counts(c, M, M, len);
    Buf<double> augf(c, M), U(c, M);
    SSN_LAUNCH(c, aug_fill_kernel, cdiv(M, 256), 256, 0, nc, N, r.p, perm.p, blocks.p, QK.p, f.p, Ae.ptr.p, Ae.idx.p, Ae.val.p,
               Aug.ptr.p, Aug.idx.p, Aug.val.p, augf.p);
    ssn_pcg_options po{};
    po.retol = opts ? opts->retol : -1.0; po.maxit = opts ? opts->maxit : -1;
    po.precd = 2; po.nf = 0; po.guess_dev = nullptr;                          // aug_PCG.m:29,32
    pcg_solve(c, Aug, augf, &po, U, itpcg, respcg, nullptr);
    SSN_LAUNCH(c, aug_recombine_kernel, cdiv(N, 256), 256, 0, nc, N, blocks.p, U.p, qp.p, zeta);
    SSN_CUDA(cudaStreamSynchronize(c->stream));
    if (info) { info[0] = nc; info[1] = 1; }
}

// =================================================================== POT bordering

namespace {

__global__ void mask_mul_kernel(int64_t n, const uint8_t* __restrict__ s, const double* __restrict__ phi, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = s[i] ? phi[i] : 0.0;
}
__global__ void pot_w_kernel(int N, const double* __restrict__ z, double coef, const double* __restrict__ v, double* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) w[i] = z[i] - coef * v[i];
}
__global__ void pot_zeta_kernel(int N, const double* __restrict__ ww, const double* __restrict__ vv, double coef, double* __restrict__ zeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) zeta[i] = ww[i] + coef * vv[i];
}

struct PotPrologue { Buf<double> v, w; double z2, sg, phi_e; };

void pot_prologue(ssn_ctx* c, const ssn_prob_data* pd, PotPrologue& P) {
    SSN_REQUIRE(pd && pd->s_dev && pd->phi_dev && pd->z_dev, SSN_E_INVALID, "POT prob_data needs s, phi and z");
    const int64_t m = pd->m, n = pd->n, mn = m * n; const int N = (int)(m + n);
    P.sg = 1.0 / pd->tk;
    Buf<double> sphi(c, mn);
    SSN_LAUNCH(c, mask_mul_kernel, 148 * 8, 256, 0, mn, pd->s_dev, pd->phi_dev, sphi.p);
    P.phi_e = pd->bk1 + P.sg * dev_dot(c, pd->phi_dev, sphi, mn);            // AMG4POT.m:33
    P.v.alloc(c, N); P.w.alloc(c, N);
    plan_ax(c, sphi, pd->p_dev, pd->q_dev, m, n, P.v);                        // AMG4POT.m:34
    P.z2 = read_scalar(c, pd->z_dev + N);
    SSN_LAUNCH(c, pot_w_kernel, cdiv(N, 256), 256, 0, N, pd->z_dev, P.sg / P.phi_e * P.z2, P.v.p, P.w.p);
}

void pot_epilogue(ssn_ctx* c, int N, PotPrologue& P, const double* vv, const double* ww, double* zeta) {
    const double vvv = dev_dot(c, P.v, vv, N), vww = dev_dot(c, P.v, ww, N);
    const double tt = P.sg * P.sg / (P.phi_e - P.sg * P.sg * vvv);           // AMG4POT.m:53
    SSN_LAUNCH(c, pot_zeta_kernel, cdiv(N, 256), 256, 0, N, ww, vv, tt * vww, zeta);
    const double vz = dev_dot(c, P.v, zeta, N);
    const double zeta2 = (P.z2 - P.sg * vz) / P.phi_e;
    upload_small(c, zeta + N, &zeta2, sizeof(double));
    SSN_CUDA(cudaStreamSynchronize(c->stream));
}

}  // namespace

void amg4pot(ssn_ctx* c, const ssn_prob_data* pd, const ssn_amg_options* opts, double* zeta, int* it, double* res, int* info, bool twogrid) {
    PotPrologue P; pot_prologue(c, pd, P);
    const int N = (int)(pd->m + pd->n);
    Buf<double> vv(c, N), ww(c, N);
    ssn_prob_data q = *pd;
    int it1, it2, i1[2], i2[2]; double r1, r2;
    q.z_dev = P.v; hybrid_amg(c, &q, opts, vv, &it1, &r1, i1, twogrid);      // AMG4POT.m:46 / :49 (str = 'twogrid': Hybrid_twogrid)
    q.z_dev = P.w; hybrid_amg(c, &q, opts, ww, &it2, &r2, i2, twogrid);      // AMG4POT.m:47 / :50
    pot_epilogue(c, N, P, vv, ww, zeta);
    if (it) *it = std::max(it1, it2);
    if (res) *res = std::max(r1, r2);
    if (info) { info[0] = std::max(i1[0], i2[0]); info[1] = std::max(i1[1], i2[1]); }
}

void pcg4pot(ssn_ctx* c, const ssn_prob_data* pd, const ssn_pcg_options* opts, double* zeta, int* it, double* res, int* info) {
    PotPrologue P; pot_prologue(c, pd, P);
    const int N = (int)(pd->m + pd->n);
    Buf<double> vv(c, N), ww(c, N);
    ssn_prob_data q = *pd;
    int it1, it2, i1[2], i2[2]; double r1, r2;
    q.z_dev = P.v; aug_pcg(c, &q, opts, vv, &it1, &r1, i1);                  // PCG4POT.m:35
    q.z_dev = P.w; aug_pcg(c, &q, opts, ww, &it2, &r2, i2);                  // PCG4POT.m:36
    pot_epilogue(c, N, P, vv, ww, zeta);
    if (it) *it = std::max(it1, it2);
    if (res) *res = std::max(r1, r2);
    if (info) { info[0] = std::max(i1[0], i2[0]); info[1] = std::max(i1[1], i2[1]); }
}

// =================================================================== invAAt / invHHt

namespace {

__global__ void invaat_kernel(int n, int m, const double* __restrict__ x, const double* __restrict__ p, const double* __restrict__ q,
                              double sg1, double sg2, double np_, double nq, double qvn, double pvm, double* __restrict__ y) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n + m) return;
    const double den = sg1 * sg2 + sg1 * nq + sg2 * np_;
    if (v < n) y[v] = x[v] / (sg1 + np_) + (np_ / (sg1 + np_) * qvn - pvm) * q[v] / den;           // invAAt.m:17
    else       y[v] = x[v] / (sg2 + nq) + (nq / (sg2 + nq) * pvm - qvn) * p[v - n] / den;           // invAAt.m:18
}
__global__ void invhht_kernel(int N, double s, double lVv1, double v2, const double* __restrict__ Vv1, const double* __restrict__ Vl,
                              double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) y[i] = (s * Vv1[i] + lVv1 * Vl[i] - v2 * Vl[i]) / s;                                 // invHHt.m:14,17
    else if (i == N) y[N] = (v2 - lVv1) / s;                                                        // invHHt.m:15,17
}

}  // namespace

void invaat(ssn_ctx* c, const double* x, const double* p, const double* q, int64_t m, int64_t n, double sg1, double sg2, double* y) {
    SSN_REQUIRE(x && p && q && y && m > 0 && n > 0, SSN_E_INVALID, "invAAt: bad arguments");
    const double np_ = dev_dot(c, p, p, m), nq = dev_dot(c, q, q, n);
    const double qvn = dev_dot(c, q, x, n), pvm = dev_dot(c, p, x + n, m);
    SSN_LAUNCH(c, invaat_kernel, cdiv(m + n, 256), 256, 0, (int)n, (int)m, x, p, q, sg1, sg2, np_, nq, qvn, pvm, y);
}

void invhht(ssn_ctx* c, const double* v, const double* p, const double* q, int64_t m, int64_t n, double sg, const double* phi, double* y) {
    SSN_REQUIRE(v && p && q && y && phi && m > 0 && n > 0, SSN_E_INVALID, "invHHt: bad arguments");
    const int N = (int)(m + n);
    const double t = sg + dev_dot(c, phi, phi, m * n);                       // invHHt.m:8
    Buf<double> l(c, N), Vl(c, N), Vv1(c, N);
    plan_ax(c, phi, p, q, m, n, l);
    invaat(c, l, p, q, m, n, sg + 1.0, sg + 1.0, Vl);                        // invHHt.m:9
    const double s = t - dev_dot(c, l, Vl, N);
    invaat(c, v, p, q, m, n, sg + 1.0, sg + 1.0, Vv1);                       // invHHt.m:12
    const double lVv1 = dev_dot(c, l, Vv1, N);
    const double v2 = read_scalar(c, v + N);
    SSN_LAUNCH(c, invhht_kernel, cdiv(N + 1, 256), 256, 0, N, s, lVv1, v2, Vv1.p, Vl.p, y);
}

}  // namespace ssn
