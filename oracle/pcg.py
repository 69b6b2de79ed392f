"""Oracle: preconditioned CG -- reference PCG.m (test infrastructure only)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _default_options(e):
    return {"guess": np.zeros_like(e), "retol": 1e-11, "maxit": 1e4, "precd": 2}


def _is_empty(v):
    return v is None or (hasattr(v, "__len__") and len(v) == 0)


def ichol0(H):
    """``ichol(H)`` with MATLAB's default options (type 'nofill', no diagonal compensation): the lower-triangular
    IC(0) factor ``L`` with the pattern of ``tril(H)``, ``L_ij = (h_ij - sum_{t<j} L_it L_jt) / L_jj``,
    ``L_ii = sqrt(h_ii - sum_t L_it^2)``, rows in order.  MATLAB's ``ichol`` is a built-in (its summation order
    is not documented); a nonpositive pivot is its error 'Encountered nonpositive pivot'."""
    A = sp.csr_matrix(sp.tril(H)); A.sort_indices()
    n = A.shape[0]
    indptr, indices, data = A.indptr, A.indices, A.data
    Lp = [0]; Li = []; Lv = []
    for i in range(n):
        r0 = len(Li)
        has_diag = False
        for e in range(indptr[i], indptr[i + 1]):
            j = int(indices[e])
            s = float(data[e])
            a0, a1 = r0, len(Li)
            b0, b1 = (Lp[j], Lp[j + 1] - 1) if j < i else (r0, a1)
            while a0 < a1 and b0 < b1:
                if Li[a0] == Li[b0]:
                    s -= Lv[a0] * Lv[b0]; a0 += 1; b0 += 1
                elif Li[a0] < Li[b0]:
                    a0 += 1
                else:
                    b0 += 1
            if j < i:
                Li.append(j); Lv.append(s / Lv[Lp[j + 1] - 1])
            else:
                if not s > 0.0:
                    raise ValueError("Encountered nonpositive pivot.")
                Li.append(i); Lv.append(float(np.sqrt(s))); has_diag = True
        if not has_diag:
            raise ValueError("Encountered nonpositive pivot.")
        Lp.append(len(Li))
    return sp.csr_matrix((np.array(Lv), np.array(Li, dtype=np.int64), np.array(Lp, dtype=np.int64)), shape=(n, n))


def PCG(H, e, pcg_options=None):
    """``[d,it,res,resk] = PCG(H,e[,pcg_options])`` -- reference PCG.m:1-105.

    Preconditioners (PCG.m:34-66): 1 none, 2 Jacobi, 3 SSOR(w=1.5), 5 bi-SSOR for the bigraph
    block form (needs ``nf``).  4 (``ichol``) is not restated (no demo selects it).
    The stopping test is on ``r'M^{-1}r`` against ``tol^2 * delta_0`` (PCG.m:76); a zero rhs
    skips the loop and returns ``res = NaN`` (PCG.m:87, 0/0).
    """
    e = np.asarray(e, dtype=np.float64).reshape(-1)
    if pcg_options is None:                                   # PCG.m:18-23
        pcg_options = _default_options(e)
    opts = dict(pcg_options)
    if _is_empty(opts.get("guess")): opts["guess"] = np.zeros_like(e)      # PCG.m:24-27
    if _is_empty(opts.get("retol")): opts["retol"] = 1e-11
    if _is_empty(opts.get("maxit")): opts["maxit"] = 1e4
    if _is_empty(opts.get("precd")): opts["precd"] = 2
    ii = int(opts["precd"])
    d0 = np.asarray(opts["guess"], dtype=np.float64).reshape(-1)
    tol = float(opts["retol"])
    maxit = int(opts["maxit"])
    Hs = H.tocsc() if sp.issparse(H) else np.asarray(H, dtype=np.float64)

    if ii == 1:
        P = None
    elif ii == 2:
        P = np.asarray(Hs.diagonal()).reshape(-1)             # PCG.m:38
    elif ii == 3:                                             # PCG.m:40-41
        D = sp.diags(Hs.diagonal()).tocsc()
        P = (D, sp.tril(Hs, -1).tocsc(), sp.triu(Hs, 1).tocsc())
    elif ii == 4:                                             # PCG.m:45-51
        if not sp.issparse(H):
            raise ValueError("iC requires H is sparse!")      # PCG.m:50
        P = ichol0(Hs)
    elif ii == 5:                                             # PCG.m:55-62
        if "nf" not in opts:
            raise ValueError("SSOR for bigraph requires pcg_options.nf!!!")      # PCG.m:64
        w = 1.5; Nf = int(opts["nf"]); Hc = sp.csc_matrix(Hs)
        V = Hc[:Nf, :Nf]; U = Hc[:Nf, Nf:]; T = Hc[Nf:, Nf:]
        invV = sp.diags(1.0 / V.diagonal()); invT = sp.diags(1.0 / T.diagonal())
        P = w * (2 - w) * sp.bmat([[invV + w ** 2 * invV @ U @ invT @ U.T @ invV, -w * invV @ U @ invT],
                                    [-w * invT @ U.T @ invV, invT]]).tocsc()
    else:
        raise ValueError("unknown precd")

    def pre(r):                                               # PCG.m:90-105
        if ii == 1:
            return r.copy()
        if ii == 2:
            return r / P
        if ii == 3:
            w = 1.5
            p1 = spla.spsolve_triangular((P[0] + w * P[1]).tocsr(), r, lower=True)
            p2 = P[0] @ p1
            # PCG.m:99: `w*(2-w) * (D+wU) \ p2` parses as ((w*(2-w))*(D+wU)) \ p2
            return spla.spsolve_triangular((w * (2 - w) * (P[0] + w * P[2])).tocsr(), p2, lower=False)
        if ii == 4:                                           # PCG.m:100-101  p = P\r; p = P'\p
            p1 = spla.spsolve_triangular(P, r, lower=True)
            return spla.spsolve_triangular(P.T.tocsr(), p1, lower=False)
        return P @ r

    it = 0
    r = e - Hs @ d0                                           # PCG.m:68
    p = pre(r)
    delta_new = float(r @ p); d = d0.copy()
    delta_0 = delta_new
    resk = np.zeros(maxit)
    while it < maxit and delta_new > tol ** 2 * delta_0:      # PCG.m:76
        delta_old = delta_new
        q = Hs @ p
        alpha = delta_old / float(q @ p)
        d = d + alpha * p
        r = r - alpha * q
        w_ = pre(r)
        delta_new = float(r @ w_)
        beta = delta_new / delta_old
        p = w_ + beta * p
        it += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            resk[it - 1] = np.sqrt(abs(delta_new / delta_0))
    with np.errstate(invalid="ignore", divide="ignore"):
        res = float(np.sqrt(abs(np.float64(delta_new) / np.float64(delta_0))))   # PCG.m:87
    return d, it, res, resk
