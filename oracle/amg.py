"""Oracle: classical AMG setup and cycles -- reference AMG/*.m (test infrastructure only).

All sparse matrices are ``scipy.sparse.csc_matrix`` with sorted indices and no explicit zeros
(what a MATLAB sparse matrix holds).  Sparse products use the frozen Gustavson order of
``oracle/csrc/oracle_kernels.c``; the random stream is ``oracle.rng.GLOBAL_STREAM``.
"""
import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import _ck
from . import rng as _rng
from .pcg import PCG


# ---------------------------------------------------------------- sparse helpers

def _csc(A):
    A = sp.csc_matrix(A, dtype=np.float64)
    A.eliminate_zeros()
    A.sort_indices()
    return A


def spgemm(A, B):
    """``A*B`` in the frozen order: C(:,j) = sum_k A(:,k)*B(k,j), k ascending, no FMA."""
    A = _csc(A); B = _csc(B)
    m, k = A.shape
    k2, n = B.shape
    assert k == k2
    Ap = A.indptr.astype(np.int64); Ai = A.indices.astype(np.int64)
    Bp = B.indptr.astype(np.int64); Bi = B.indices.astype(np.int64)
    Cp = np.zeros(n + 1, dtype=np.int64)
    L = _ck.lib()
    nnz = L.orc_spgemm_csc_symbolic(m, n, Ap, Ai, Bp, Bi, Cp)
    Ci = np.zeros(max(nnz, 1), dtype=np.int64)
    Cx = np.zeros(max(nnz, 1), dtype=np.float64)
    L.orc_spgemm_csc_numeric(m, n, Ap, Ai, np.ascontiguousarray(A.data), Bp, Bi,
                             np.ascontiguousarray(B.data), Cp, Ci, Cx)
    C = sp.csc_matrix((Cx[:nnz], Ci[:nnz], Cp), shape=(m, n))
    C.eliminate_zeros()                       # MATLAB sparse results hold no explicit zeros
    C.sort_indices()
    return C


def spmv(A, x):
    """``A*x`` for sparse A in CSC column-sweep order (y(i) += A(i,j)*x(j), j ascending)."""
    A = _csc(A)
    y = np.empty(A.shape[0], dtype=np.float64)
    _ck.lib().orc_spmv_csc(A.shape[0], A.shape[1], A.indptr.astype(np.int64),
                           A.indices.astype(np.int64), np.ascontiguousarray(A.data),
                           np.ascontiguousarray(np.asarray(x, dtype=np.float64)), y)
    return y


def _diag_solve(D, B):
    """``D\\B`` for diagonal D: element-wise division of row i by d_i (frozen convention)."""
    d = np.asarray(D, dtype=np.float64).reshape(-1)
    B = _csc(B).tocoo()
    with np.errstate(divide="ignore", invalid="ignore"):
        vals = B.data / d[B.row]
    return _csc(sp.csc_matrix((vals, (B.row, B.col)), shape=B.shape))


# ---------------------------------------------------------------- strength / splitting

def strength(A, which=2):
    """Strength values ``S`` -- reference AMG/strength.m:6-18."""
    A = _csc(A)
    N = A.shape[0]
    A0 = _csc(sp.diags(A.diagonal(), format="csc") - A)        # strength.m:7-8
    coo = A0.tocoo()                                           # find(): column-major order
    ia, ja, s0 = coo.row, coo.col, coo.data
    max_row = np.zeros(N)                                      # max incl. implicit zeros: the
    if s0.size:                                                # diagonal of A0 is always one
        np.maximum.at(max_row, ia, s0)
    max_row[max_row <= 0] = np.inf                             # strength.m:10
    with np.errstate(divide="ignore", invalid="ignore"):
        if which == 1:
            sa = s0 / max_row[ia]
        else:
            sa = s0 / np.minimum(max_row[ia], max_row[ja])     # strength.m:16
    return _csc(sp.csc_matrix((sa, (ia, ja)), shape=(N, N)))


def _logical(S, theta):
    """``S >= theta`` as a sparse logical (theta > 0: only stored entries can be true)."""
    S = _csc(S).tocoo()
    keep = S.data >= theta
    As = sp.csc_matrix((np.ones(int(keep.sum())), (S.row[keep], S.col[keep])), shape=S.shape)
    As.sort_indices()
    return As


def mis_set(A, theta=0.025):
    """iFEM-style randomised MIS C/F split -- reference AMG/mis_set.m:8-67.

    Returns ``(isC, isF, As)``; consumes the global MATLAB random stream
    (mis_set.m:31 N0 draws on the degenerate branch, else mis_set.m:35 one draw per node with
    deg>0, ascending node order).
    """
    A = _csc(A)
    N = A.shape[0]
    isF = np.zeros(N, dtype=bool)
    isC = np.zeros(N, dtype=bool)
    N0 = min(int(math.sqrt(N)) + 1, 25)                        # mis_set.m:12
    As = _logical(strength(A), theta)                          # mis_set.m:25
    deg = np.asarray(As.sum(axis=0)).reshape(-1).astype(np.float64)   # column counts, :28-29
    if np.count_nonzero(deg > 0) < 0.25 * math.sqrt(N):        # mis_set.m:30-34
        pick = np.ceil(_rng.rand(N0) * N).astype(np.int64) - 1
        isC[pick] = True
        isF = ~isC
        return isC, isF, As
    idx = deg > 0
    deg[idx] = deg[idx] + 0.1 * _rng.rand(int(idx.sum()))      # mis_set.m:35
    isF[deg == 0] = True                                       # mis_set.m:40
    isU = np.ones(N, dtype=bool)
    Ac = As.tocoo()
    er, ec = Ac.row, Ac.col
    upper = er < ec                                            # triu(.,1) entries
    ur, uc = er[upper], ec[upper]
    while isC.sum() < N / 2 and isU.sum() > N0:                # mis_set.m:42
        isS = deg > 0                                          # :44-45
        live = isS[ur] & isS[uc]                               # edges of As(S,S), i<j
        i_, j_ = ur[live], uc[live]
        ge = deg[i_] >= deg[j_]                                # :50
        kill = np.concatenate([j_[ge], i_[~ge]])               # :51-52
        isS[kill] = False
        isC[isS] = True                                        # :53
        hit = isC[ec]                                          # find(As(:,isC)) rows, :56-57
        isF[er[hit]] = True
        isU = ~(isF | isC)
        deg[~isU] = 0                                          # :59
        if isU.sum() <= N0:                                    # :61-64
            isC[isU] = True
            isU = np.zeros(0, dtype=bool)
    iso = np.asarray(As.sum(axis=1)).reshape(-1) == 0          # mis_set.m:67
    isC[iso] = True
    isF[iso] = False
    return isC, isF, As


def cf_split(S):
    """Sequential greedy C/F split in index order -- reference AMG/cf_split.m:6-15.

    ``S`` is the logical strength matrix; MATLAB ``graph(S)`` needs it symmetric and the
    neighbours of k are the off-diagonal nonzeros of row/column k.  The third output of the
    reference (a MATLAB ``graph`` object) is returned here as the CSR adjacency.
    """
    G = sp.csr_matrix(S).astype(bool).astype(np.int8)
    G.setdiag(0); G.eliminate_zeros(); G.sort_indices()
    N = G.shape[0]
    indF = np.zeros(N, dtype=bool); indC = np.zeros(N, dtype=bool); indU = np.ones(N, dtype=bool)
    for k in range(N):
        if indU[k]:
            kk = G.indices[G.indptr[k]:G.indptr[k + 1]]
            indC[k] = True; indU[k] = False
            indF[kk] = True; indU[kk] = False
    return indC, indF, G


# ---------------------------------------------------------------- hierarchy state

class _State:
    """The reference's globals ``Ack Prok J smoth_it Rk`` (AMG/Class_AMG.m:42-43)."""

    def __init__(self):
        self.clear()
        self.last = None          # snapshot of the last hierarchy, for tests

    def clear(self):
        self.Ack = []; self.Prok = []; self.Rk = []; self.J = 0; self.smoth_it = 0
        self.trace = []           # per-level (isC, As) records, for parity tests


amg_state = _State()

MAX_LEVELS = 64


class AMGError(RuntimeError):
    pass


def _opt(opts, key, default):
    v = opts.get(key, None)
    if v is None or (hasattr(v, "__len__") and not isinstance(v, str) and len(v) == 0):
        return default
    return v


def transfer(A, amg_options=None, J=None, want_aux=False):
    """``[Ac,Pro,As,indC] = transfer(A[,amg_options])`` -- reference AMG/transfer.m:8-66.

    ``J`` stands for the reference's ``global J`` (transfer.m:17); default: the oracle's
    hierarchy state.  Returns ``(Ac, Pro)`` or ``(Ac, Pro, As, indC)`` with ``want_aux``.
    """
    if amg_options is None:                                    # transfer.m:8-10
        amg_options = {"theta": 1 / 40, "bigph": 0, "inter": 1, "diag": 0}
    theta = _opt(amg_options, "theta", 1 / 4)                  # transfer.m:12-15
    bigph = _opt(amg_options, "bigph", 0)
    inter = _opt(amg_options, "inter", 1)
    isnsp = _opt(amg_options, "isnsp", 0)
    if J is None:
        J = amg_state.J
    A = _csc(A)
    N = A.shape[0]
    As = None; indC = None
    if J == 1 and bigph:                                       # transfer.m:19-29
        Nf = int(amg_options["fnode"]); Nc = N - Nf
        Aff = A[:Nf, :Nf]; Afc = _csc(A[:Nf, Nf:])
        offd = _csc(Aff - sp.diags(Aff.diagonal()))
        if offd.nnz == 0:
            W = _diag_solve(-Aff.diagonal(), Afc)              # W = (-Aff)\Afc, diagonal Aff
        else:                                                  # non-bigraph block: general solve
            W = _csc(spla.spsolve(_csc(-Aff), Afc))
        if isnsp == 1:
            W = _diag_solve(spmv(W, np.ones(Nc)), W)           # transfer.m:23
        Pro = _csc(sp.vstack([W, sp.identity(Nc, format="csc")]))
        if want_aux:
            indC = np.zeros(N, dtype=bool); indC[Nf:] = True
            As = _logical(strength(A), theta)
    else:
        indC, indF, As = mis_set(A, theta)                     # transfer.m:41
        amg_state.trace.append({"isC": indC.copy(), "isF": indF.copy(), "As": As.copy()})
        C_node = np.flatnonzero(indC); Nc = C_node.size
        F_node = np.flatnonzero(indF); Nf = F_node.size
        if Nc + Nf != N or np.any(indC & indF):
            raise AMGError("C/F split does not partition the nodes (AMG/transfer.m:46 would "
                           "index out of range)")
        p = np.concatenate([F_node, C_node])
        AA = _csc(A[p, :][:, p])                               # transfer.m:46
        Aff = _csc(AA[:Nf, :Nf]); Afc = _csc(AA[:Nf, Nf:])
        if inter < 2:
            dff = Aff.diagonal()
            W1 = _diag_solve(-dff, Afc)                        # transfer.m:49  (-Dff)\Afc
            as_ = _csc(sp.identity(Nf, format="csc") + As[F_node, :][:, F_node])   # :50
            Affs = _csc(Aff.multiply(as_))                     # :51  Aff.*as
            W2 = spgemm(_diag_solve(-dff, Affs), W1)           # :51  ((-Dff)\Affs)*W1
            W = _csc(W1 + inter * W2)                          # :52
            if Nf > 0:                                         # :54 `~isempty(...)`: always
                W = _csc(W1 + 0.5 * W2)                        # :55
        else:
            W = _csc(spla.spsolve(_csc(-Aff), Afc))            # :58
        if isnsp == 1:
            W = _diag_solve(spmv(W, np.ones(Nc)), W)           # :60-62
        P = _csc(sp.vstack([W, sp.identity(Nc, format="csc")]))
        inv = np.empty(N, dtype=np.int64); inv[p] = np.arange(N)
        Pro = _csc(P[inv, :])                                  # Pro(p,:) = P   (:63)
    Ac = spgemm(spgemm(_csc(Pro.T), A), Pro)                   # transfer.m:66
    if want_aux:
        return Ac, Pro, As, indC
    return Ac, Pro


# ---------------------------------------------------------------- cycles

def _smooth(A, R, r, e, isnsp, n_it, aux):
    if isnsp:                                                  # MG_Wcycle.m:15-21
        xx, Axi = aux
        for _ in range(n_it):
            g = r - A @ e
            xig = g.sum()
            g = (xig / xx) + R @ (g - Axi * (xig / xx))
            e = e + g
    else:                                                      # MG_Wcycle.m:23
        for _ in range(n_it):
            e = e + R @ (r - A @ e)
    return e


def _kernel_aux(A):
    N = A.shape[0]
    xi = np.ones(N)
    Axi = A @ xi
    xx = float(xi @ Axi)                                       # xi'*A*xi
    return xx, Axi


def MG_Vcycle(r, isnsp=0, k=1):
    """Recursive V-cycle -- reference AMG/MG_Vcycle.m:5-45 (``k`` is 1-based)."""
    st = amg_state
    R = st.Rk[k - 1]; A = st.Ack[k - 1]; Rt = R.T.tocsc()
    r = np.asarray(r, dtype=np.float64)
    if k < st.J:
        aux = _kernel_aux(A) if isnsp else None
        e = _smooth(A, R, r, np.zeros_like(r), isnsp, st.smoth_it, aux)
        Pn = st.Prok[k]
        rrc = Pn.T @ (r - A @ e)
        eec = MG_Vcycle(rrc, isnsp, k + 1)
        e = e + Pn @ eec
        e = _smooth(A, Rt, r, e, isnsp, st.smoth_it, aux)
    else:
        e, _, _, _ = PCG(A, r)                                 # MG_Vcycle.m:43
    return e


def MG_Wcycle(r, isnsp=0, k=1, e=None):
    """Recursive W-cycle -- reference AMG/MG_Wcycle.m:5-46 (``k`` is 1-based)."""
    st = amg_state
    R = st.Rk[k - 1]; A = st.Ack[k - 1]; Rt = R.T.tocsc()
    r = np.asarray(r, dtype=np.float64)
    if e is None:
        e = np.zeros_like(r)
    if k < st.J:
        aux = _kernel_aux(A) if isnsp else None
        e = _smooth(A, R, r, e, isnsp, st.smoth_it, aux)
        Pn = st.Prok[k]
        rrc = Pn.T @ (r - A @ e)
        eec = MG_Wcycle(rrc, isnsp, k + 1)                     # MG_Wcycle.m:28
        eec = MG_Wcycle(rrc, isnsp, k + 1, eec)                # MG_Wcycle.m:30
        e = e + Pn @ eec
        e = _smooth(A, Rt, r, e, isnsp, st.smoth_it, aux)
    else:
        e, _, _, _ = PCG(A, r)                                 # MG_Wcycle.m:44 (guess ignored)
    return e


# ---------------------------------------------------------------- Class_AMG

def coarsest_size_threshold(N):
    """``1 + fix(N^(1/3))`` with the host libm ``pow`` -- AMG/Class_AMG.m:76."""
    return 1 + int(math.pow(float(N), 1.0 / 3.0))


def setup_hierarchy(A, amg_options):
    """Setup phase of Class_AMG -- reference AMG/Class_AMG.m:41-85 (fills ``amg_state``)."""
    st = amg_state
    st.clear()
    st.smoth_it = int(amg_options["smoth"])
    A = _csc(A)
    st.J = 1; Ak = A
    st.Ack.append(Ak); st.Prok.append(None)
    dofk = A.shape[0]
    if amg_options["bigph"]:                                   # Class_AMG.m:48-59
        Nf = int(amg_options["fnode"]); Nc = dofk - Nf
        V = Ak[:Nf, :Nf]; U = _csc(Ak[:Nf, Nf:]); T = Ak[Nf:, Nf:]
        with np.errstate(divide="ignore"):
            invV = sp.diags(1.0 / V.diagonal(), format="csc")
            invT = sp.diags(1.0 / T.diagonal(), format="csc")
        low = _csc(-(invT @ U.T @ invV))
        st.Rk.append(_csc(sp.bmat([[invV, None], [low, invT]], format="csc")))
    else:
        with np.errstate(divide="ignore"):
            st.Rk.append(_csc(0.5 * sp.diags(1.0 / Ak.diagonal(), format="csc")))   # :72
    thr = coarsest_size_threshold(A.shape[0])
    while Ak.shape[0] > thr:                                   # Class_AMG.m:76
        if st.J >= MAX_LEVELS:
            raise AMGError("coarsening stalled")
        Ak_new, Pro = transfer(st.Ack[st.J - 1], amg_options, J=st.J)
        if Ak_new.shape[0] >= Ak.shape[0]:
            raise AMGError("coarsening stalled (no F nodes); MATLAB would loop forever")
        Ak = Ak_new
        st.J += 1; st.Ack.append(Ak); st.Prok.append(Pro)
        with np.errstate(divide="ignore"):
            st.Rk.append(_csc(0.5 * sp.diags(1.0 / Ak.diagonal(), format="csc")))   # :84
    return st


def Class_AMG(A, b, amg_options=None):
    """``[x,it,rel_res,rel_resk,rhok] = Class_AMG(A,b[,opts])`` -- AMG/Class_AMG.m:20-110."""
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    if amg_options is None:                                    # Class_AMG.m:20-25
        amg_options = {"retol": 1e-12, "bigph": 0, "maxit": 20, "theta": 1 / 4, "smoth": 10,
                       "cycle": 1, "isnsp": 1, "inter": 1, "guess": np.zeros_like(b)}
    o = dict(amg_options)
    o["retol"] = _opt(o, "retol", 1e-12); o["bigph"] = _opt(o, "bigph", 0)     # :26-34
    o["maxit"] = int(_opt(o, "maxit", 50)); o["theta"] = _opt(o, "theta", 1 / 4)
    o["smoth"] = int(_opt(o, "smoth", 3)); o["cycle"] = _opt(o, "cycle", "v")
    o["isnsp"] = _opt(o, "isnsp", 0); o["inter"] = _opt(o, "inter", 1)
    o["guess"] = np.asarray(_opt(o, "guess", np.zeros_like(b)), dtype=np.float64).reshape(-1)
    if o["bigph"]:
        if _opt(o, "fnode", None) is None or o["fnode"] <= 0:
            raise AMGError("amg_options.bigph = 1 requires Nf > 0")           # :36-40
    A = _csc(A)
    st = setup_hierarchy(A, o)
    it = 0
    maxit = o["maxit"]
    rhok = [np.nan]
    rel_resk = [1.0]
    x = o["guess"].copy()
    res0 = np.linalg.norm(A @ x - b)                           # Class_AMG.m:89
    rel_res = None
    if res0 == 0:
        rel_res = 0.0; rel_resk = np.array([0.0]); rhok = np.array([np.inf])
    else:
        it = 1
        while rel_resk[it - 1] > o["retol"] and it <= maxit:   # :95
            r = b - A @ x
            if o["cycle"] == "v":
                x = x + MG_Vcycle(r, o["isnsp"])
            if o["cycle"] == "w":
                x = x + MG_Wcycle(r, o["isnsp"])
            res = np.linalg.norm(A @ x - b); rel_res = res / res0
            rel_resk.append(rel_res)
            with np.errstate(divide="ignore", invalid="ignore"):
                rhok.append(res / np.linalg.norm(r))
            it += 1
            if rhok[it - 1] > 1:
                break
        rel_resk = np.array(rel_resk[:it]); rhok = np.array(rhok[:it]); it -= 1
    st.last = {"Ack": list(st.Ack), "Prok": list(st.Prok), "Rk": list(st.Rk), "J": st.J,
               "trace": list(st.trace)}
    st.clear()                                                 # Class_AMG.m:110
    return x, it, rel_res, rel_resk, rhok


# ---------------------------------------------------------------- two-grid (Hybrid_twogrid, inner_solver = 5)

def twogrid_bigph(A, b, amg_options=None):
    """``[x,it,rel_res,rel_resk,rhok] = twogrid_bigph(A,b[,amg_options])`` -- AMG/twogrid_bigph.m:1-116.

    Setup (:26-47): block Gauss-Seidel smoother ``R = [invV 0; -invT*Afc'*invV invT]``, interpolation
    ``W = -Aff\\Afc`` (diagonal ``Aff``), row-normalised when ``isnsp``, ``Ac = Pro'*A*Pro``.  One iteration
    (``twogrid_it``, :79-116): ``smoth`` pre-smoothing steps with ``R``, restriction, coarse correction
    ``PCG(Ac, rrc, retol [] -> 1e-11, maxit 100, Jacobi)``, prolongation, ``smoth`` post-smoothing steps with ``R'``."""
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    if amg_options is None:                                    # :14-18
        amg_options = {"retol": 1e-12, "maxit": 20, "fnode": 0, "smoth": 10, "isnsp": 1, "guess": np.zeros_like(b)}
    o = dict(amg_options)
    o["retol"] = _opt(o, "retol", 0.0); o["maxit"] = int(_opt(o, "maxit", 50))         # :19-23
    o["smoth"] = int(_opt(o, "smoth", 3)); o["isnsp"] = _opt(o, "isnsp", 0)
    o["guess"] = np.asarray(_opt(o, "guess", np.zeros_like(b)), dtype=np.float64).reshape(-1)
    A = _csc(A)
    N = A.shape[0]
    Nf = int(_opt(o, "fnode", 0)); Nc = N - Nf                 # :28
    if not (0 < Nf < N):
        raise AMGError("twogrid_bigph requires 0 < amg_options.fnode < N")
    Aff = A[:Nf, :Nf]; Afc = _csc(A[:Nf, Nf:]); Acc = A[Nf:, Nf:]
    with np.errstate(divide="ignore"):
        invV = sp.diags(1.0 / Aff.diagonal(), format="csc")
        invT = sp.diags(1.0 / Acc.diagonal(), format="csc")
    R = _csc(sp.bmat([[invV, None], [_csc(-(invT @ Afc.T @ invV)), invT]], format="csc"))      # :35
    W = _diag_solve(-Aff.diagonal(), Afc)                      # :42  W = -Aff\Afc (Aff diagonal on a bigraph)
    if o["isnsp"] == 1:
        W = _diag_solve(spmv(W, np.ones(Nc)), W)               # :44
    Pro = _csc(sp.vstack([W, sp.identity(Nc, format="csc")]))
    Ac = spgemm(spgemm(_csc(Pro.T), A), Pro)                   # :47
    return _twogrid_solve(A, b, o, R, Pro, Ac)


def _twogrid_solve(A, b, o, R, Pro, Ac):
    """Solve phase shared by twogrid_bigph.m:57-116 and twogrid.m:95-150 (identical text)."""
    Rt = _csc(R.T)
    aux = _kernel_aux(A) if o["isnsp"] else None
    pcg_options = {"retol": None, "maxit": 100, "precd": 2, "guess": None}              # twogrid_bigph.m:98, twogrid.m:136

    def twogrid_it(r):
        e = _smooth(A, R, r, np.zeros_like(r), o["isnsp"], o["smoth"], aux)            # pre-smoothing
        rrc = Pro.T @ (r - A @ e)                              # restriction
        eec = PCG(Ac, rrc, pcg_options)[0]                     # coarse correction
        e = e + Pro @ eec                                      # prolongation
        return _smooth(A, Rt, r, e, o["isnsp"], o["smoth"], aux)                       # post-smoothing with R'

    it = 0
    rhok = [np.nan]; rel_resk = [1.0]
    x = o["guess"].copy()
    res0 = np.linalg.norm(A @ x - b)
    rel_res = None
    if res0 == 0:
        rel_res = 0.0; rel_resk = np.array([0.0]); rhok = np.array([np.inf])
    else:
        it = 1
        while rel_resk[it - 1] > o["retol"] and it <= o["maxit"]:
            r = b - A @ x
            x = x + twogrid_it(r)
            res = np.linalg.norm(A @ x - b); rel_res = res / res0
            rel_resk.append(rel_res)
            with np.errstate(divide="ignore", invalid="ignore"):
                rhok.append(res / np.linalg.norm(r))
            it += 1
            if rhok[it - 1] > 1:
                break
        rel_resk = np.array(rel_resk[:it]); rhok = np.array(rhok[:it]); it -= 1
    return x, it, rel_res, rel_resk, rhok


def twogrid(A, b, amg_options=None):
    """``[x,it,rel_res,rel_resk,rhok] = twogrid(A,b[,amg_options])`` -- AMG/twogrid.m:1-150: the two-grid method
    for a general graph Laplacian.  ``bigph = 1``: the bigraph setup of twogrid_bigph after the checks of
    :36-38 and :46-48; ``bigph = 0``: damped Jacobi ``0.5*D^-1`` (:59) and ``mis_set(A,1/4)`` + the standard
    interpolation ``W1 + 0.5*W2`` (:73-92, the MIS step of transfer.m with theta = 1/4)."""
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    if amg_options is None:                                    # :16-21
        amg_options = {"retol": 1e-12, "bigph": 0, "maxit": 20, "smoth": 10, "isnsp": 1, "guess": np.zeros_like(b)}
    o = dict(amg_options)
    o["retol"] = _opt(o, "retol", 0.0); o["bigph"] = _opt(o, "bigph", 0)               # :22-34
    o["maxit"] = int(_opt(o, "maxit", 50)); o["smoth"] = int(_opt(o, "smoth", 3)); o["isnsp"] = _opt(o, "isnsp", 0)
    o["guess"] = np.asarray(_opt(o, "guess", np.zeros_like(b)), dtype=np.float64).reshape(-1)
    o["fnode"] = int(_opt(o, "fnode", 0))
    if o["bigph"] and o["fnode"] <= 0:
        raise AMGError("bigph = 1 requires fnode > 0")                                  # :36-38
    A = _csc(A)
    if o["bigph"]:
        Nf = o["fnode"]
        Aff = _csc(A[:Nf, :Nf])
        if _csc(Aff - sp.diags(Aff.diagonal())).nnz and abs(Aff - sp.diags(Aff.diagonal())).max() != 0:
            raise AMGError("Nf is not right for the bigraph")                           # :46-48
        return twogrid_bigph(A, b, o)
    with np.errstate(divide="ignore"):
        R = _csc(0.5 * sp.diags(1.0 / A.diagonal(), format="csc"))                      # :59
    Ac, Pro = transfer(A, {"theta": 1 / 4, "bigph": 0, "inter": 1, "isnsp": o["isnsp"]}, J=2)       # :73-94
    return _twogrid_solve(A, b, o, R, Pro, Ac)
