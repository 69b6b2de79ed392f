"""Oracle: matrix-free plan operators and assembly (test infrastructure only).

Conventions: ``x = vec(X)`` is the column-major m x n plan; dual vectors
are ``[column part (n) ; row part (m)]``; ``p`` has length m, ``q`` length n.
"""
import numpy as np
import scipy.sparse as sp


def _col(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))


def Ax(x, p, q):
    """``y = [X'p ; Xq]`` -- reference Ax.m:10-13."""
    p = _col(p); q = _col(q)
    m, n = p.size, q.size
    if sp.issparse(x):                       # Class1/warmup_class1.m:29 passes a sparse zero x
        x = np.asarray(x.todense())
    X = _col(x).reshape((m, n), order="F")
    r = X.T @ p
    l = X @ q
    return np.concatenate([r, l])


def Aty(y, p, q):
    """``z = vec(p*y1' + y2*q')`` -- reference Aty.m:10-13."""
    p = _col(p); q = _col(q); y = _col(y)
    m, n = p.size, q.size
    y1, y2 = y[:n], y[n:n + m]
    Z = np.multiply.outer(p, y1) + np.multiply.outer(y2, q)     # Aty.m:12-13 (mul, mul, add)
    return Z.reshape(-1, order="F")


def explicit_A(p, q):
    """``A = [kron(speye(n),p'); kron(q',speye(m))]`` -- Class1/APD_SsN_Class1.m:47 (comment)."""
    p = _col(p); q = _col(q)
    m, n = p.size, q.size
    top = sp.kron(sp.identity(n, format="csr"), sp.csr_matrix(p.reshape(1, m)))
    bot = sp.kron(sp.csr_matrix(q.reshape(1, n)), sp.identity(m, format="csr"))
    return sp.vstack([top, bot]).tocsr()


def active_pattern(s, m, n):
    """(rows, cols) of the nonzeros of ``Y = sparse(reshape(s,m,n))`` in CSC order -- ASAt.m:15."""
    s = np.asarray(s).reshape(-1)
    lin = np.flatnonzero(s)
    return (lin % m).astype(np.int64), (lin // m).astype(np.int64)


def ASAt(s, p, q):
    """Sparse ``H = A*diag(s)*A'`` -- reference ASAt.m:14-19.

    Returns CSC, sorted row indices, explicit zeros dropped (what MATLAB ``sparse`` holds).
    Block order: column nodes 0..n-1, then row nodes n..n+m-1.  Diagonal sums are accumulated
    sequentially in ascending index order, starting from 0.0 (CSC mat-vec order).
    """
    p = _col(p); q = _col(q)
    m, n = p.size, q.size
    ii, jj = active_pattern(s, m, n)                     # CSC order: j ascending, i ascending
    # U = P*Y -> U_ij = p_i ; Q = Y*R -> Q_ij = q_j                    (ASAt.m:18)
    dcol = np.bincount(jj, weights=p[ii] * p[ii], minlength=n)   # U'*p  (ascending i per j)
    drow = np.bincount(ii, weights=q[jj] * q[jj], minlength=m)   # Q*q   (ascending j per i)
    ru = q[jj] * p[ii]                                   # R*U' entry (j, i)
    pq = p[ii] * q[jj]                                   # P*Q  entry (i, j)
    N = n + m
    rows = np.concatenate([np.arange(n), jj, n + ii, n + np.arange(m)])
    cols = np.concatenate([np.arange(n), n + ii, jj, n + np.arange(m)])
    vals = np.concatenate([dcol, ru, pq, drow])
    H = sp.csc_matrix((vals, (rows, cols)), shape=(N, N))
    H.eliminate_zeros()
    H.sort_indices()
    return H


def ASAtz(z, s, p, q):
    """Matrix-free ``A*diag(s)*A'*z`` -- reference ASAtz.m:15-22.

    Reproduces the reference as written, including ``Q*p`` at ASAtz.m:21 (where ASAt.m:19 uses
    ``Q*q``): only dimensionally valid for m == n and only equal to ``ASAt*z`` when p == q.
    """
    p = _col(p); q = _col(q); z = _col(z)
    m, n = p.size, q.size
    if m != n:
        raise ValueError("ASAtz.m:21 multiplies the m-by-n matrix Q by p (length m): needs m == n")
    Y = np.asarray(s, dtype=np.float64).reshape((m, n), order="F")
    U = p[:, None] * Y
    Q = Y * q[None, :]
    z1, z2 = z[:n], z[n:n + m]
    y1 = (U.T @ p) * z1 + q * (U.T @ z2)
    y2 = p * (Q @ z1) + (Q @ p) * z2
    return np.concatenate([y1, y2])


def invAAt(x, p, q, sg1=None, sg2=None):
    """``(diag(sg1*I_n, sg2*I_m) + A*A') \\ x`` in closed form -- reference invAAt.m:7-20."""
    if sg1 is None:
        sg1, sg2 = 1.0, 1.0
    elif sg2 is None:
        sg2 = sg1
    p = _col(p); q = _col(q); x = _col(x)
    m, n = p.size, q.size
    np_ = np.linalg.norm(p) ** 2
    nq = np.linalg.norm(q) ** 2
    vn, vm = x[:n], x[n:n + m]
    den = sg1 * sg2 + sg1 * nq + sg2 * np_
    yn = vn / (sg1 + np_) + (np_ / (sg1 + np_) * (q @ vn) - p @ vm) * q / den
    ym = vm / (sg2 + nq) + (nq / (sg2 + nq) * (p @ vm) - q @ vn) * p / den
    return np.concatenate([yn, ym])


def invHHt(v, p, q, sg, phi):
    """``(sg*I + H*H') \\ v`` with ``H = (G,IY,IZ)``, ``G = [A;phi']`` -- Class2/invHHt.m:7-17."""
    p = _col(p); q = _col(q); v = _col(v); phi = _col(phi)
    m, n = p.size, q.size
    t = sg + np.linalg.norm(phi) ** 2
    l = Ax(phi, p, q)
    Vl = invAAt(l, p, q, sg + 1)
    s = t - l @ Vl
    v1, v2 = v[:n + m], v[-1]
    Vv1 = invAAt(v1, p, q, sg + 1)
    y1 = s * Vv1 + (l @ Vv1) * Vl - v2 * Vl
    y2 = v2 - l @ Vv1
    return np.concatenate([y1, [y2]]) / s
