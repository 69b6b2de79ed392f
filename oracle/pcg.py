"""Oracle: preconditioned CG -- reference PCG.m (test infrastructure only)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _default_options(e):
    return {"guess": np.zeros_like(e), "retol": 1e-11, "maxit": 1e4, "precd": 2}


def _is_empty(v):
    return v is None or (hasattr(v, "__len__") and len(v) == 0)


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
    elif ii == 4:
        raise NotImplementedError("precd=4 (ichol, PCG.m:46) is not restated; SURVEY.md 8f row 3")
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
