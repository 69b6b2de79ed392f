"""Oracle: inner linear-solve dispatch -- reference Hybrid_AMG.m, aug_PCG.m, components.m,
Class2/AMG4POT.m, Class2/PCG4POT.m (test infrastructure only)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.csgraph as csgraph
import scipy.sparse.linalg as spla

from . import rng as _rng
from .amg import Class_AMG, twogrid_bigph, _csc
from .pcg import PCG
from .plan_ops import Ax


def components(A):
    """``[blocks,sizes,p,r] = components(A)`` -- reference components.m:32-55.

    FROZEN CONVENTION (deviation from MATLAB, whose ``dmperm`` ordering is not available):
    components are numbered by ascending smallest member and ``p`` lists the members of each
    component in ascending order.  ``blocks`` is 1-based (labels 1..k); ``p`` holds 0-based
    node indices; ``r`` holds 0-based block boundaries (component i is ``p[r[i]:r[i+1]]``).
    """
    A = sp.csr_matrix(A)
    n, m = A.shape
    if n != m:
        raise ValueError("Adjacency matrix must be square")              # components.m:33
    pat = sp.csr_matrix((np.ones(A.nnz, dtype=np.int8), A.indices, A.indptr), shape=A.shape)
    pat.data[np.asarray(A.data) == 0] = 0
    pat.eliminate_zeros()
    k, lab = csgraph.connected_components(pat, directed=False)
    first = np.full(k, n, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(n))
    order = np.argsort(first, kind="stable")
    rank = np.empty(k, dtype=np.int64); rank[order] = np.arange(k)
    blocks0 = rank[lab]
    p = np.argsort(blocks0, kind="stable")
    sizes = np.bincount(blocks0, minlength=k)
    r = np.concatenate([[0], np.cumsum(sizes)])
    return blocks0 + 1, sizes, p, r


def rescaled_system(prob_data):
    """``Q0, A0, Q, K, Ae, f`` -- reference Hybrid_AMG.m:12-24 (== aug_PCG.m:11-22).

    Entry-wise rounding is pinned: ``A0_ij = (qp_i*h_ij)*qp_j`` (``(Q0*H0)*Q0``),
    ``Q_ii = qp_i*qp_i``, ``K_ii = (qp_i*t_i)*qp_i``, ``Ae = bk1*Q + (1/tk)*(K+A0)``.
    """
    bk1 = float(prob_data["bk1"]); tk = float(prob_data["tk"])
    q = np.asarray(prob_data["q"], dtype=np.float64).reshape(-1)
    p = np.asarray(prob_data["p"], dtype=np.float64).reshape(-1)
    H0 = _csc(prob_data["H0"]); z = np.asarray(prob_data["z"], dtype=np.float64).reshape(-1)
    T = prob_data["T"]
    t = np.asarray(T.diagonal() if sp.issparse(T) else np.asarray(T)).reshape(-1).astype(np.float64)
    M = H0.shape[0]
    qp = np.concatenate([q, -p])
    if np.any(qp == 0):
        raise ValueError("p or q contains 0 !!!!!")                      # Hybrid_AMG.m:18-19
    coo = H0.tocoo()
    A0 = _csc(sp.csc_matrix(((qp[coo.row] * coo.data) * qp[coo.col], (coo.row, coo.col)), shape=(M, M)))
    Qd = qp * qp
    Kd = (qp * t) * qp
    inv_tk = 1.0 / tk
    KA = _csc(sp.diags(Kd, format="csc") + A0)
    Ae = _csc(bk1 * sp.diags(Qd, format="csc") + inv_tk * KA)
    f = qp * z
    return qp, A0, Qd, Kd, Ae, f


def Hybrid_twogrid(prob_data, amg_options):
    """``[zeta,itamg,resamg,info] = Hybrid_twogrid(prob_data,amg_options)`` -- Hybrid_twogrid.m:11-89: the
    dispatch of Hybrid_AMG with ``twogrid_bigph`` (:39, :67) in place of ``Class_AMG``."""
    return Hybrid_AMG(prob_data, amg_options, _solver=twogrid_bigph)


def Hybrid_AMG(prob_data, amg_options, _solver=None):
    """``[zeta,itamg,resamg,info] = Hybrid_AMG(prob_data,amg_options)`` -- Hybrid_AMG.m:11-113."""
    solver = _solver or Class_AMG
    bk1 = float(prob_data["bk1"]); tk = float(prob_data["tk"])
    q = np.asarray(prob_data["q"]).reshape(-1)
    qp, A0, Qd, Kd, Ae, f = rescaled_system(prob_data)
    M = A0.shape[0]
    blocks, sizes, ps, rs = components(A0)                               # Hybrid_AMG.m:27
    num_comp = len(sizes)
    n = q.size
    opts = dict(amg_options)
    u = np.zeros(M)
    itamg = 0; resamg = 0.0; it_num = 0
    if num_comp == 1:                                                    # :30-48
        opts["isnsp"] = 0 if Kd.sum() else 1
        opts["fnode"] = n
        opts["guess"] = (bk1 * tk) * _rng.rand(M)                        # :40
        u, itamg, resamg, _, _ = solver(Ae, f, opts)
        it_num = 1
    if num_comp > 1:                                                     # :50-107
        N0 = 100
        large = np.flatnonzero(sizes > N0)
        for k in large:
            pk = ps[rs[k]:rs[k + 1]]
            Aek = _csc(Ae[pk, :][:, pk]); fk = f[pk]
            opts["isnsp"] = 0 if Kd[pk].sum() else 1
            opts["fnode"] = int(np.count_nonzero(pk < n))                # :68  sum(pk<=n), 1-based
            opts["guess"] = (bk1 * tk) * _rng.rand(pk.size)              # :69
            dk, itk, resk, _, _ = solver(Aek, fk, opts)
            u[pk] = dk; itamg = max(itamg, itk); resamg = max(resamg, resk)
            it_num = int(k) + 1                                          # :80 (1-based k)
        b2s = sizes[blocks - 1]
        small = np.flatnonzero(b2s <= N0)
        bb = blocks[small]
        pk = small[np.argsort(bb, kind="stable")]
        if pk.size:                                                      # :87-91
            A0s = _csc(A0[pk, :][:, pk])
            Aes = _csc(bk1 * sp.diags(Qd[pk], format="csc")
                       + (1.0 / tk) * _csc(sp.diags(Kd[pk], format="csc") + A0s))
            u[pk] = spla.spsolve(Aes, f[pk]) if pk.size > 1 else f[pk] / Aes[0, 0]
    zeta = qp * u                                                        # :113
    return zeta, itamg, resamg, np.array([num_comp, it_num])


def aug_PCG(prob_data, pcg_options):
    """``[zeta,itpcg,respcg,info] = aug_PCG(prob_data,pcg_options)`` -- aug_PCG.m:11-37."""
    bk1 = float(prob_data["bk1"]); tk = float(prob_data["tk"])
    qp, A0, Qd, Kd, Ae, f = rescaled_system(prob_data)
    M = A0.shape[0]
    blocks, sizes, _, _ = components(A0)                                 # aug_PCG.m:24
    nc = len(sizes)
    Y = sp.csc_matrix((np.ones(M), (np.arange(M), blocks - 1)), shape=(M, nc))
    QK = sp.diags(bk1 * Qd + (1.0 / tk) * Kd, format="csc")              # aug_PCG.m:27
    augAe = sp.bmat([[Y.T @ QK @ Y, Y.T @ QK], [QK @ Y, Ae]], format="csc")
    augf = np.concatenate([Y.T @ f, f])
    o = dict(pcg_options)
    o["guess"] = np.zeros(nc + M)
    o["precd"] = 2                                                       # aug_PCG.m:32
    U, itpcg, respcg, _ = PCG(augAe, augf, o)
    u = Y @ U[:nc] + U[nc:]
    return qp * u, itpcg, respcg, np.array([nc, 1])


def _pot_split(prob_data):
    """Common prologue of AMG4POT.m:27-34 / PCG4POT.m:26-33."""
    p = np.asarray(prob_data["p"]).reshape(-1); q = np.asarray(prob_data["q"]).reshape(-1)
    bk1 = float(prob_data["bk1"]); tk = float(prob_data["tk"])
    phi = np.asarray(prob_data["phi"], dtype=np.float64).reshape(-1)
    z = np.asarray(prob_data["z"], dtype=np.float64).reshape(-1)
    s = np.asarray(prob_data["s"]).reshape(-1).astype(np.float64)
    z1, z2 = z[:-1], z[-1]
    epss, sg = bk1, 1.0 / tk
    phi_e = epss + sg * (phi @ (s * phi))
    v = Ax(s * phi, p, q)
    w = z1 - sg / phi_e * z2 * v
    return z2, sg, phi_e, v, w


def _pot_combine(z2, sg, phi_e, v, vv, ww):
    tt = sg ** 2 / (phi_e - sg ** 2 * (v @ vv))                          # AMG4POT.m:53
    zeta1 = ww + tt * vv * (v @ ww)
    zeta2 = (z2 - sg * (v @ zeta1)) / phi_e
    return np.concatenate([zeta1, [zeta2]])


def AMG4POT(prob_data, amg_options, str_="amg"):
    """Bordered POT solve by two Hybrid_AMG calls (``str = 'amg'``) or two Hybrid_twogrid calls (any other ``str``, e.g.
    ``'twogrid'``) -- reference Class2/AMG4POT.m:27-55."""
    solve = Hybrid_AMG if str_ == "amg" else Hybrid_twogrid             # AMG4POT.m:45-51
    z2, sg, phi_e, v, w = _pot_split(prob_data)
    pd = dict(prob_data)
    pd["z"] = v; vv, it1, res1, info1 = solve(pd, amg_options)           # AMG4POT.m:46 / :49
    pd["z"] = w; ww, it2, res2, info2 = solve(pd, amg_options)           # AMG4POT.m:47 / :50
    zeta = _pot_combine(z2, sg, phi_e, v, vv, ww)
    return zeta, max(it1, it2), max(res1, res2), np.maximum(info1, info2)


def PCG4POT(prob_data, pcg_options):
    """Bordered POT solve by two aug_PCG calls -- reference Class2/PCG4POT.m:26-40."""
    z2, sg, phi_e, v, w = _pot_split(prob_data)
    pd = dict(prob_data)
    pd["z"] = v; vv, it1, res1, info1 = aug_PCG(pd, pcg_options)
    pd["z"] = w; ww, it2, res2, info2 = aug_PCG(pd, pcg_options)
    zeta = _pot_combine(z2, sg, phi_e, v, vv, ww)
    return zeta, max(it1, it2), max(res1, res2), np.maximum(info1, info2)
