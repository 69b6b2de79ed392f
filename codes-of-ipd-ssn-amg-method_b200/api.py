"""Host-side mirror of the reference's MATLAB functions on the SsN inner-solve path.

Same names, same positional arguments and the same error behaviour as the reference ``.m``
files; every function is a thin marshalling layer over one C-ABI entry
point of ``libssnamg.so`` -- this file stands where the MEX shims stand for MATLAB.

Array arguments may be NumPy arrays (host; copied to the device, results come back as NumPy)
or torch CUDA tensors (device-resident, the analogue of ``gpuArray``; results are torch CUDA
tensors).  Sparse matrices are returned as :class:`DeviceCSR` handles (device-resident CSR; use
``.to_scipy()``); functions taking a sparse matrix accept a :class:`DeviceCSR` or any SciPy
sparse matrix.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import AmgOptions, CSR, PcgOptions, ProbData, SsnError, context


def _torch():
    import torch
    return torch


# ------------------------------------------------------------------ marshalling helpers

def _is_host(*xs):
    torch = _torch()
    return not any(isinstance(x, torch.Tensor) for x in xs if x is not None)


def _dev(x, dtype=None, count=None):
    """-> contiguous 1-D CUDA tensor of the given dtype."""
    torch = _torch()
    dtype = dtype or torch.float64
    if isinstance(x, torch.Tensor):
        t = x
        if t.dtype != dtype:
            t = t.to(dtype)
        if not t.is_cuda:
            t = t.cuda()
        if t.dim() == 2:
            t = t.t().reshape(-1)                  # column-major, like MATLAB's X(:) and like the NumPy branch below
        elif t.dim() != 1:
            raise ValueError("expected a vector or a matrix (flattened column-major)")
        t = t.contiguous()
    else:
        a = np.asarray(x)
        npd = {torch.float64: np.float64, torch.uint8: np.uint8, torch.int32: np.int32, torch.int64: np.int64}[dtype]
        if a.dtype == np.bool_ and npd is np.uint8:
            a = a.view(np.uint8)
        a = np.ascontiguousarray(a.reshape(-1, order="F"), dtype=npd)
        t = torch.from_numpy(a).cuda()
    if count is not None and t.numel() != count:
        raise ValueError(f"expected {count} elements, got {t.numel()}")
    return t


def _ret(t, host):
    return t.cpu().numpy() if host else t


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DeviceCSR:
    """Device-resident CSR matrix handle (``ssn_csr``).  Structurally symmetric matrices on this
    path have CSR arrays == the CSC arrays a MATLAB sparse matrix holds."""

    def __init__(self, ctx, st, owned=True, keepalive=None):
        self.ctx, self.st, self.owned, self._keep = ctx, st, owned, keepalive

    @property
    def shape(self):
        return (int(self.st.nrows), int(self.st.ncols))

    @property
    def nnz(self):
        return int(self.st.nnz)

    @classmethod
    def from_scipy(cls, A, ctx=None):
        import scipy.sparse as sp
        ctx = ctx or context()
        A = sp.csr_matrix(A, dtype=np.float64)
        A.sort_indices()
        st = CSR()
        rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        v = np.ascontiguousarray(A.data, dtype=np.float64)
        ctx.call("ssn_csr_upload", A.shape[0], A.shape[1], A.nnz, rp.ctypes.data_as(C.c_void_p),
                 ci.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.byref(st))
        return cls(ctx, st)

    def to_scipy(self):
        import scipy.sparse as sp
        n, nnz = int(self.st.nrows), int(self.st.nnz)
        rp = np.empty(n + 1, dtype=np.int32); ci = np.empty(max(nnz, 1), dtype=np.int32)
        v = np.empty(max(nnz, 1), dtype=np.float64)
        self.ctx.call("ssn_csr_download", C.byref(self.st), rp.ctypes.data_as(C.c_void_p),
                      ci.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        return sp.csr_matrix((v[:nnz], ci[:nnz], rp), shape=self.shape)

    def free(self):
        if self.owned and self.st.rowptr_dev:
            try:
                self.ctx.call("ssn_csr_free", C.byref(self.st))
            except Exception:
                pass
        self.owned = False

    def __del__(self):
        self.free()


def _csr(A, ctx):
    return A if isinstance(A, DeviceCSR) else DeviceCSR.from_scipy(A, ctx)


def _empty(v):
    return v is None or (hasattr(v, "__len__") and not isinstance(v, str) and len(v) == 0)


def _amg_options(opts, keep):
    """dict (MATLAB struct) -> ssn_amg_options; missing/empty field -> 'empty' sentinel."""
    o = AmgOptions(retol=-1.0, bigph=-1, maxit=-1, theta=-1.0, smoth=-1, cycle=-1, isnsp=-1, inter=-1,
                   fnode=0, guess_dev=None)
    if opts is None:
        return None
    g = lambda k: None if _empty(opts.get(k)) else opts.get(k)
    if g("retol") is not None: o.retol = float(g("retol"))
    if g("bigph") is not None: o.bigph = int(g("bigph"))
    if g("maxit") is not None: o.maxit = int(g("maxit"))
    if g("theta") is not None: o.theta = float(g("theta"))
    if g("smoth") is not None: o.smoth = int(g("smoth"))
    if g("cycle") is not None:
        cyc = g("cycle")
        o.cycle = ord(cyc) if isinstance(cyc, str) else int(cyc)
    if g("isnsp") is not None: o.isnsp = int(g("isnsp"))
    if g("inter") is not None: o.inter = int(g("inter"))
    if g("fnode") is not None: o.fnode = int(g("fnode"))
    if g("guess") is not None:
        t = _dev(g("guess")); keep.append(t); o.guess_dev = t.data_ptr()
    return o


def _pcg_options(opts, keep):
    if opts is None:
        return None
    o = PcgOptions(retol=-1.0, maxit=-1, precd=-1, nf=0, guess_dev=None)
    g = lambda k: None if _empty(opts.get(k)) else opts.get(k)
    if g("retol") is not None: o.retol = float(g("retol"))
    if g("maxit") is not None: o.maxit = int(g("maxit"))
    if g("precd") is not None: o.precd = int(g("precd"))
    if g("nf") is not None: o.nf = int(g("nf"))
    if g("guess") is not None:
        t = _dev(g("guess")); keep.append(t); o.guess_dev = t.data_ptr()
    return o


def _byref_or_null(o):
    return C.byref(o) if o is not None else None


# ------------------------------------------------------------------ random stream

def rng_reset(seed=5489):
    """Reset the library-owned MATLAB ``rand`` stream (mt19937ar seed 0 == init_genrand(5489))."""
    context().call("ssn_rng_reset", int(seed))


def rng_drawn():
    ctx = context()
    return int(ctx.lib.ssn_rng_drawn(ctx.h))


def rand(count):
    torch = _torch()
    out = torch.empty(int(count), dtype=torch.float64, device="cuda")
    context().call("ssn_rand", int(count), _ptr(out))
    return out


# ------------------------------------------------------------------ L1: plan operators

def Ax(x, p, q):
    """``y = Ax(x,p,q)`` -- reference Ax.m:2-14."""
    torch = _torch(); ctx = context(); host = _is_host(x, p, q)
    if hasattr(x, "todense"):                      # Class1/warmup_class1.m:29 passes a sparse zero x
        x = np.asarray(x.todense())
    pd, qd = _dev(p), _dev(q)
    m, n = pd.numel(), qd.numel()
    xd = _dev(x, count=m * n)
    y = torch.empty(n + m, dtype=torch.float64, device="cuda")
    ctx.call("ssn_ax", _ptr(xd), _ptr(pd), _ptr(qd), m, n, _ptr(y))
    return _ret(y, host)


def Aty(y, p, q):
    """``z = Aty(y,p,q)`` -- reference Aty.m:2-14."""
    torch = _torch(); ctx = context(); host = _is_host(y, p, q)
    pd, qd = _dev(p), _dev(q)
    m, n = pd.numel(), qd.numel()
    yd = _dev(y)
    if yd.numel() < n + m:
        raise ValueError("y must have at least n+m entries")
    z = torch.empty(m * n, dtype=torch.float64, device="cuda")
    ctx.call("ssn_aty", _ptr(yd), _ptr(pd), _ptr(qd), m, n, _ptr(z))
    return _ret(z, host)


def prox_residual(w, lam, p, q, tk, gama=np.inf, want=("Axprox", "norm2", "count"), scal_dev=None):
    """Fused SsN residual pieces in one read of ``w`` (Class1/APD_SsN_Class1.m:139-144,184):
    ``z=(w-Aty(lam))/tk``, ``s=(z>=0)&(z<=gama)``, ``prox=min(max(0,z),gama)``, ``Ax(prox)``,
    ``||prox||^2``, ``nnz(s)``.  ``want`` selects the outputs; returns a dict."""
    torch = _torch(); ctx = context(); host = _is_host(w, lam, p, q)
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    wd, ld = _dev(w, count=m * n), _dev(lam)
    gvec, gs = None, float("inf")
    if np.isscalar(gama) or (hasattr(gama, "numel") and gama.numel() == 1) or np.size(gama) == 1:
        gs = float(gama if np.isscalar(gama) else np.asarray(gama.cpu() if hasattr(gama, "cpu") else gama).reshape(-1)[0])
    else:
        gvec = _dev(gama, count=m * n)
    mk = lambda name, shape, dt: torch.empty(shape, dtype=dt, device="cuda") if name in want else None
    axp = mk("Axprox", n + m, torch.float64); px = mk("prox", m * n, torch.float64)
    z = mk("z", m * n, torch.float64); s = mk("s", m * n, torch.uint8)
    n2 = C.c_double(0.0); cnt = C.c_int64(0)
    if scal_dev is not None:
        # ``scal_dev``: 2 device doubles that receive the norm term and nnz(s); no host read, nothing synchronised
        ctx.call("ssn_prox_residual_dev", _ptr(wd), _ptr(ld), _ptr(pd), _ptr(qd), m, n, float(tk), _ptr(gvec), gs,
                 _ptr(axp), _ptr(px), _ptr(z), _ptr(s), _ptr(scal_dev))
        out = {}
    else:
        ctx.call("ssn_prox_residual", _ptr(wd), _ptr(ld), _ptr(pd), _ptr(qd), m, n, float(tk), _ptr(gvec), gs,
                 _ptr(axp), _ptr(px), _ptr(z), _ptr(s), C.byref(n2), C.byref(cnt))
        out = {"norm2": n2.value, "count": cnt.value}
    for k, v in (("Axprox", axp), ("prox", px), ("z", z), ("s", s)):
        if v is not None:
            out[k] = _ret(v, host)
    return out


def prox_residual_pot(w, lam, p, q, tk, phi, want=("Hprox", "norm2", "count")):
    """Fused SsN residual pieces of PARTIAL OT in one read of ``w`` and one of ``phi`` (Class2/APD_SsN_Class2.m:124-130,
    137-150, 196-217; ``u = [x (m*n); y (n); z (m)]``, ``lam`` of n+m+1 entries):
    ``zk = (wk - [Aty(lam[:N]) + lam[N]*phi ; lam[:N]])/tk``, ``s = zk[:mn] >= 0`` (uint8), ``t = zk[mn:] >= 0`` (0/1 doubles),
    ``prox = max(zk, 0)``, ``Hprox = [Ax(prox x) + [prox y; prox z] ; phi'prox x]``, ``||prox||^2``, ``nnz(s)``."""
    torch = _torch(); ctx = context(); host = _is_host(w, lam, p, q)
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel(); N = m + n
    wd, ld, fd = _dev(w, count=m * n + N), _dev(lam, count=N + 1), _dev(phi, count=m * n)
    mk = lambda name, shape, dt: torch.empty(shape, dtype=dt, device="cuda") if name in want else None
    hp = mk("Hprox", N + 1, torch.float64); px = mk("prox", m * n + N, torch.float64)
    s = mk("s", m * n, torch.uint8); t = mk("t", N, torch.float64)
    n2 = C.c_double(0.0); cnt = C.c_int64(0)
    ctx.call("ssn_prox_residual_pot", _ptr(wd), _ptr(ld), _ptr(pd), _ptr(qd), m, n, float(tk), _ptr(fd),
             _ptr(hp), _ptr(px), _ptr(s), _ptr(t), C.byref(n2), C.byref(cnt))
    out = {"norm2": n2.value, "count": cnt.value}
    for k, v in (("Hprox", hp), ("prox", px), ("s", s), ("t", t)):
        if v is not None:
            out[k] = _ret(v, host)
    return out


def _gama_args(gama, m, n):
    if np.isscalar(gama) or (hasattr(gama, "numel") and gama.numel() == 1) or np.size(gama) == 1:
        gs = float(gama if np.isscalar(gama) else np.asarray(gama.cpu() if hasattr(gama, "cpu") else gama).reshape(-1)[0])
        return None, gs
    return _dev(gama, count=m * n), float("inf")


def prox_trials(w, lamT, p, q, tk, gama=np.inf):
    """``||prox((w - Aty(lam_t))/tk)||^2`` for up to 8 trial dual vectors (rows of ``lamT``) in one read
    of ``w`` -- the objective of the Armijo trials of Class1/APD_SsN_Class1.m:193-207.  Returns a
    device tensor of ``len(lamT)`` squared norms."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    wd = _dev(w, count=m * n)
    lt = lamT if isinstance(lamT, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(lamT, dtype=np.float64))).cuda()
    lt = lt.to(device="cuda", dtype=torch.float64).reshape(-1, n + m).contiguous()
    gvec, gs = _gama_args(gama, m, n)
    out = torch.empty(lt.shape[0], dtype=torch.float64, device="cuda")
    ctx.call("ssn_prox_trials", _ptr(wd), _ptr(lt), int(lt.shape[0]), _ptr(pd), _ptr(qd), m, n, float(tk), _ptr(gvec), gs, _ptr(out))
    return out


def prox_trials_lin(w, lam, zeta, p, q, tk, delta, ll0, nt):
    """``||prox((w - Aty(lam + delta**(ll0+t)*zeta))/tk)||^2`` for ``t < nt <= 256`` backtracking steps of one
    search direction in one read of ``w`` (``gama = Inf``), through the screened kernels: the values of
    ``prox_trials`` on the same trial vectors up to the summation order, HBM-bound for any ``nt`` where the
    trial plans are sparse.
    Returns a device tensor of ``nt + 1`` doubles: the squared norms and the number of entries that
    survived the screen (out of ``m*n``)."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    wd, ld, zd = _dev(w, count=m * n), _dev(lam, count=m + n), _dev(zeta, count=m + n)
    out = torch.empty(int(nt) + 1, dtype=torch.float64, device="cuda")
    ctx.call("ssn_prox_trials_lin", _ptr(wd), _ptr(ld), _ptr(zd), _ptr(pd), _ptr(qd), m, n, float(tk), float(delta), int(ll0), int(nt), _ptr(out))
    return out


def APD_SsN_Class1(c, r, l, p, q, gama=np.inf, inner_solver=4, maxit=100, KKT_Tol=1e-6, warm_maxit=100, max_outer=None,
                   max_seconds=None, verbose=False, amg_options=None, pcg_options=None, host_call=False):
    """The reference's Class 1 script (Class1/APD_SsN_Class1.m:32-275 + Class1/warmup_class1.m) as ONE call into the
    library (``ssn_apd_ssn_class1``): warm start, APD outer loop, SsN inner loop, line search, KKT bookkeeping, with no
    Python between the kernels.  ``host_call=True`` goes through ``ssn_apd_ssn_class1_host`` with NumPy arrays (inputs
    copied to the device once, plan and duals copied back once).  Returns the dictionary of ``driver.APD_SsN_Class1``."""
    from ._lib import ApdOptions, ApdResult
    torch = _torch(); ctx = context()
    keep = []
    ao = _amg_options(amg_options, keep); po = _pcg_options(pcg_options, keep)
    o = ApdOptions(inner_solver=int(inner_solver), maxit=int(maxit), KKT_Tol=float(KKT_Tol), warm_maxit=int(warm_maxit),
                   max_outer=int(max_outer or 0), max_seconds=float(max_seconds or 0.0), verbose=1 if verbose else 0,
                   amg=C.pointer(ao) if ao is not None else None, pcg=C.pointer(po) if po is not None else None)
    res = ApdResult()
    hist = np.zeros((3, int(maxit) + 1)); its = np.zeros(int(maxit), dtype=np.int32)
    cap = 64 * int(maxit)
    steps = np.zeros((cap, 7))
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    if host_call:
        f = lambda a: np.ascontiguousarray(np.asarray(a.cpu() if hasattr(a, "cpu") else a, dtype=np.float64).reshape(-1))
        ch, rh, lh, ph, qh = f(c), f(r), f(l), f(p), f(q); m, n = ph.size, qh.size
        scalar = np.isscalar(gama) or np.size(gama) == 1
        gh = None if scalar else f(gama)
        xk = np.empty(m * n); lk = np.empty(m + n)
        ctx.call("ssn_apd_ssn_class1_host", hp(ch), hp(rh), hp(lh), hp(ph), hp(qh), m, n, hp(gh) if gh is not None else None,
                 float(gama) if scalar else float("inf"), C.byref(o), hp(xk), hp(lk), C.byref(res), hp(hist[0]), hp(hist[1]), hp(hist[2]),
                 hp(its), hp(steps), cap)
    else:
        pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
        cd = _dev(c, count=m * n); rd = _dev(r, count=n); ld = _dev(l, count=m)
        gvec, gs = _gama_args(gama, m, n)
        xk = torch.empty(m * n, dtype=torch.float64, device="cuda"); lk = torch.empty(n + m, dtype=torch.float64, device="cuda")
        ctx.call("ssn_apd_ssn_class1", _ptr(cd), _ptr(rd), _ptr(ld), _ptr(pd), _ptr(qd), m, n, _ptr(gvec), gs, C.byref(o), _ptr(xk), _ptr(lk),
                 C.byref(res), hp(hist[0]), hp(hist[1]), hp(hist[2]), hp(its), hp(steps), cap)
    L = res.hist_len; ns = min(int(res.steps_len), cap)
    stats = {"ssn_its": its[:res.outer_its].tolist(), "ls_trials": res.ls_trials, "ls_passes": res.ls_passes, "converged": bool(res.converged),
             "amg_calls": res.amg_calls, "warmup_s": res.warmup_s, "solve_s": res.solve_s, "asat_s": res.asat_s, "plan_s": res.plan_s,
             "steps": [tuple(row) for row in steps[:ns].tolist()]}
    return {"xk": xk, "lk": lk, "fxk": hist[0, :L].tolist(), "KKT_xk": hist[1, :L].tolist(), "KKT_lk": hist[2, :L].tolist(),
            "outer_its": res.outer_its, "rel_kkt": res.rel_kkt, "stats": stats, "seconds": res.loop_s, "warmup_seconds": res.warmup_s}


def warmup_class1(c, r, l, p, q, gama=np.inf, res=None, maxit=None):
    """``[xk,lk] = warmup_class1(c,r,l,p,q,gama,res,maxit)`` -- reference Class1/warmup_class1.m:2-96
    (A-ADMM warm start), device resident.  ``nargin`` rules of :3-20: ``res`` defaults to 1e-1 and
    ``maxit`` to inf, ``res == 0 and maxit == inf`` is an error, ``maxit == inf`` means 500; the
    residual test itself is commented out in the reference (:83-91), so ``maxit`` iterations run."""
    torch = _torch(); ctx = context(); host = _is_host(c, r, l, p, q)
    if res is None:
        res = 1e-1
    if maxit is None:
        maxit = np.inf
    elif res == 0 and maxit == np.inf:
        raise ValueError("res = 0 and maxit = inf")                       # warmup_class1.m:11
    if maxit == np.inf:
        maxit = 500                                                       # :19
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    cd = _dev(c, count=m * n)
    b = torch.cat([_dev(r), _dev(l)])
    gvec, gs = _gama_args(gama, m, n)
    xk = torch.empty(m * n, dtype=torch.float64, device="cuda"); lk = torch.empty(n + m, dtype=torch.float64, device="cuda")
    ctx.call("ssn_warmup_class1", _ptr(cd), _ptr(b), _ptr(pd), _ptr(qd), m, n, _ptr(gvec), gs, int(maxit), _ptr(xk), _ptr(lk))
    return _ret(xk, host), _ret(lk, host)


def warm_stage(stage, xk, vk, wk, pik, lk2, dd, c, p, q, b, lk1, axk, y, ak, bk, gk, gama=np.inf):
    """One fused stage of a warm-start iteration (Class1/warmup_class1.m:63-67 / :70-75) on CUDA tensors that
    are updated IN PLACE (``xk vk wk pik lk2 dd``: contiguous float64 CUDA tensors of ``m*n`` entries).
    Stage 0 returns ``Ax(dd)``; stage 1 returns ``(Ax(vk1), Ax(xk1))`` -- column sums over the rows held here,
    then those rows' sums.  For callers that own the loop (the row-sharded driver)."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    for t in (xk, vk, wk, pik, lk2, dd):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == m * n):
            raise ValueError("warm_stage updates contiguous float64 CUDA tensors of m*n entries in place")
    cd, bd = _dev(c, count=m * n), _dev(b, count=m + n)
    gvec, gs = _gama_args(gama, m, n)
    out1 = torch.empty(n + m, dtype=torch.float64, device="cuda")
    if stage == 0:
        l1, ax = _dev(lk1, count=m + n), _dev(axk, count=m + n)
        ctx.call("ssn_warm_stage", 0, _ptr(xk), _ptr(vk), _ptr(wk), _ptr(pik), _ptr(lk2), _ptr(dd), _ptr(cd), _ptr(pd), _ptr(qd), _ptr(bd),
                 _ptr(l1), _ptr(ax), None, m, n, _ptr(gvec), gs, float(ak), float(bk), float(gk), _ptr(out1), None)
        return out1
    yd = _dev(y, count=m + n)
    out2 = torch.empty(n + m, dtype=torch.float64, device="cuda")
    ctx.call("ssn_warm_stage", 1, _ptr(xk), _ptr(vk), _ptr(wk), _ptr(pik), _ptr(lk2), _ptr(dd), _ptr(cd), _ptr(pd), _ptr(qd), _ptr(bd),
             None, None, _ptr(yd), m, n, _ptr(gvec), gs, float(ak), float(bk), float(gk), _ptr(out1), _ptr(out2))
    return out1, out2


def apd_begin(c, xk, vk, p, q, ak, bk):
    """``wk = -c + bk*(xk+ak*vk)/ak^2`` and ``Ax(xk)`` in one pass (Class1/APD_SsN_Class1.m:125-126)."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    cd, xd, vd = _dev(c, count=m * n), _dev(xk, count=m * n), _dev(vk, count=m * n)
    wk = torch.empty(m * n, dtype=torch.float64, device="cuda"); axk = torch.empty(n + m, dtype=torch.float64, device="cuda")
    ctx.call("ssn_apd_begin", _ptr(cd), _ptr(xd), _ptr(vd), _ptr(pd), _ptr(qd), m, n, float(ak), float(bk), _ptr(wk), _ptr(axk))
    return wk, axk


def apd_end(c, wk, xk, lam, p, q, tk, ak, gama=np.inf):
    """``xk1 = prox((wk-Aty(lam))/tk)``, ``vk1 = xk1+(xk1-xk)/ak``, ``Ax(xk1)``, ``c'xk1`` and
    ``||xk1-prox(xk1-c-Aty(lam))||^2`` in one pass (Class1/APD_SsN_Class1.m:239-254)."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    cd, wd, xd, ld = _dev(c, count=m * n), _dev(wk, count=m * n), _dev(xk, count=m * n), _dev(lam)
    gvec, gs = _gama_args(gama, m, n)
    xk1 = torch.empty(m * n, dtype=torch.float64, device="cuda"); vk1 = torch.empty(m * n, dtype=torch.float64, device="cuda")
    axk1 = torch.empty(n + m, dtype=torch.float64, device="cuda")
    cx = C.c_double(0.0); kx2 = C.c_double(0.0)
    ctx.call("ssn_apd_end", _ptr(cd), _ptr(wd), _ptr(xd), _ptr(ld), _ptr(pd), _ptr(qd), m, n, float(tk), float(ak), _ptr(gvec), gs,
             _ptr(xk1), _ptr(vk1), _ptr(axk1), C.byref(cx), C.byref(kx2))
    return xk1, vk1, axk1, cx.value, kx2.value


def trial_vectors(lam, zeta, wlk, delta, ll0, nt):
    """``lamT[t] = lam + delta**(ll0+t)*zeta`` (t < nt <= 256) and ``f0[2t] = ||lamT[t]||^2, f0[2t+1] = wlk'lamT[t]``
    as device tensors -- the O(m+n) half of a batch of Armijo trials."""
    torch = _torch(); ctx = context()
    ld, zd, wd = _dev(lam), _dev(zeta), _dev(wlk)
    N = ld.numel()
    lamT = torch.empty((int(nt), N), dtype=torch.float64, device="cuda")
    f0 = torch.empty(2 * int(nt), dtype=torch.float64, device="cuda")
    ctx.call("ssn_trial_vectors", _ptr(ld), _ptr(zd), _ptr(wd), N, float(delta), int(ll0), int(nt), _ptr(lamT), _ptr(f0))
    return lamT, f0


def linesearch(w, lam_old, zeta, wlk, p, q, tk, bk1, cF_old, ress, gama=np.inf, nu=0.2, delta=0.9, ll_max=500, batch=0):
    """Armijo backtracking of Class1/APD_SsN_Class1.m:182-211 with ``batch`` (1..8) backtracking steps per
    read of ``w`` (the full step ll = 0 is tried alone first); ``batch = 0`` (default) is adaptive: 8 to 128
    steps per read through the screened kernels, by the measured sparsity of the trial plans.
    Returns ``(lk_new, ll, norm2, cF_new, passes)``."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    wd, lo, ze, wl = _dev(w, count=m * n), _dev(lam_old), _dev(zeta), _dev(wlk)
    gvec, gs = _gama_args(gama, m, n)
    out = torch.empty(n + m, dtype=torch.float64, device="cuda")
    ll = C.c_int(0); passes = C.c_int(0); n2 = C.c_double(0.0); cF = C.c_double(0.0)
    ctx.call("ssn_linesearch", _ptr(wd), _ptr(lo), _ptr(ze), _ptr(wl), _ptr(pd), _ptr(qd), m, n, float(tk), float(bk1),
             _ptr(gvec), gs, float(nu), float(delta), int(ll_max), float(cF_old), float(ress), int(batch), _ptr(out),
             C.byref(ll), C.byref(n2), C.byref(cF), C.byref(passes))
    return out, ll.value, n2.value, cF.value, passes.value


def ASAt(s, p, q):
    """``H = ASAt(s,p,q)`` -- reference ASAt.m:2-20 (``s`` logical, one byte per entry)."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    sd = _dev(s, torch.uint8, count=m * n)
    st = CSR()
    ctx.call("ssn_asat", _ptr(sd), _ptr(pd), _ptr(qd), m, n, C.byref(st))
    return DeviceCSR(ctx, st)


def active_coo(s_loc, m_loc, n, row_offset=0, m_global=None):
    """Global column-major linear indices (int64, 0-based, slab CSC order) of the active entries of a
    row slab of the plan -- the exchange format of the row-sharded ASAt."""
    torch = _torch(); ctx = context()
    sd = _dev(s_loc, torch.uint8, count=m_loc * n)
    ptr = C.c_void_p(); E = C.c_int64(0)
    ctx.call("ssn_active_coo", _ptr(sd), int(m_loc), int(n), int(row_offset), int(m_global if m_global is not None else m_loc),
             C.byref(ptr), C.byref(E))
    out = torch.empty(E.value, dtype=torch.int64, device="cuda")
    if E.value:
        ctx.check(ctx.lib.ssn_memcpy_d2d(ctx.h, C.c_void_p(out.data_ptr()), ptr, C.c_size_t(8 * E.value)))
    ctx.call("ssn_free", ptr)
    return out


def ASAt_coo(lin_sorted, p, q):
    """``H = ASAt(s,p,q)`` from ``find(s)-1`` (ascending global linear indices, int64)."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    ld = _dev(lin_sorted, torch.int64)
    st = CSR()
    ctx.call("ssn_asat_coo", _ptr(ld), int(ld.numel()), _ptr(pd), _ptr(qd), m, n, C.byref(st))
    return DeviceCSR(ctx, st)


def ASAtz(z, s, p, q):
    """``y = ASAtz(z,s,p,q)`` -- reference ASAtz.m:2-23 (as written, ``Q*p`` at :21)."""
    torch = _torch(); ctx = context(); host = _is_host(z, s, p, q)
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    zd = _dev(z, count=n + m); sd = _dev(s, torch.uint8, count=m * n)
    y = torch.empty(n + m, dtype=torch.float64, device="cuda")
    ctx.call("ssn_asatz", _ptr(zd), _ptr(sd), _ptr(pd), _ptr(qd), m, n, _ptr(y))
    return _ret(y, host)


def invAAt(x, p, q, sg1=None, sg2=None):
    """``y = invAAt(x,p,q[,sg1[,sg2]])`` -- reference invAAt.m:1-21 (nargin defaults :7-12)."""
    torch = _torch(); ctx = context(); host = _is_host(x, p, q)
    if sg1 is None:
        sg1, sg2 = 1.0, 1.0
    elif sg2 is None:
        sg2 = sg1
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    xd = _dev(x, count=n + m)
    y = torch.empty(n + m, dtype=torch.float64, device="cuda")
    ctx.call("ssn_invaat", _ptr(xd), _ptr(pd), _ptr(qd), m, n, float(sg1), float(sg2), _ptr(y))
    return _ret(y, host)


def invHHt(v, p, q, sg, phi):
    """``y = invHHt(v,p,q,sg,phi)`` -- reference Class2/invHHt.m:1-18."""
    torch = _torch(); ctx = context(); host = _is_host(v, p, q, phi)
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    vd = _dev(v, count=n + m + 1); ph = _dev(phi, count=m * n)
    y = torch.empty(n + m + 1, dtype=torch.float64, device="cuda")
    ctx.call("ssn_invhht", _ptr(vd), _ptr(pd), _ptr(qd), m, n, float(sg), _ptr(ph), _ptr(y))
    return _ret(y, host)


# ------------------------------------------------------------------ L2: AMG setup

def strength(A, which=2):
    """``S = strength(A[,which])`` -- reference AMG/strength.m:1-19."""
    ctx = context(); Ad = _csr(A, ctx); st = CSR()
    ctx.call("ssn_strength", C.byref(Ad.st), int(which), C.byref(st))
    return DeviceCSR(ctx, st)


def mis_set(A, theta=0.025):
    """``[isC,isF,As] = mis_set(A[,theta])`` -- reference AMG/mis_set.m:1-68."""
    torch = _torch(); ctx = context(); Ad = _csr(A, ctx)
    N = Ad.shape[0]
    isC = torch.empty(N, dtype=torch.uint8, device="cuda"); isF = torch.empty_like(isC)
    st = CSR()
    ctx.call("ssn_mis_set", C.byref(Ad.st), float(theta), _ptr(isC), _ptr(isF), C.byref(st))
    return isC.cpu().numpy().astype(bool), isF.cpu().numpy().astype(bool), DeviceCSR(ctx, st)


def cf_split(S):
    """``[indC,indF,SubG] = cf_split(S)`` -- reference AMG/cf_split.m:1-16.  The third output
    (a MATLAB ``graph`` object in the reference) is returned as the adjacency handle ``S``."""
    torch = _torch(); ctx = context(); Sd = _csr(S, ctx)
    N = Sd.shape[0]
    indC = torch.empty(N, dtype=torch.uint8, device="cuda"); indF = torch.empty_like(indC)
    ctx.call("ssn_cf_split", C.byref(Sd.st), _ptr(indC), _ptr(indF))
    return indC.cpu().numpy().astype(bool), indF.cpu().numpy().astype(bool), Sd


def transfer(A, amg_options=None, J=1):
    """``[Ac,Pro,As,indC] = transfer(A[,amg_options])`` -- reference AMG/transfer.m:1-67.
    ``J`` stands for the reference's ``global J`` (transfer.m:17)."""
    torch = _torch(); ctx = context(); Ad = _csr(A, ctx); keep = []
    o = _amg_options(amg_options, keep)
    N = Ad.shape[0]
    indC = torch.empty(N, dtype=torch.uint8, device="cuda")
    ac, pro, As = CSR(), CSR(), CSR()
    ctx.call("ssn_transfer", C.byref(Ad.st), _byref_or_null(o), int(J), C.byref(ac), C.byref(pro), C.byref(As), _ptr(indC))
    return DeviceCSR(ctx, ac), DeviceCSR(ctx, pro), DeviceCSR(ctx, As), indC.cpu().numpy().astype(bool)


def amg_setup(A, amg_options):
    """Setup phase of Class_AMG (AMG/Class_AMG.m:41-85) into the library-owned hierarchy handle
    (the reference's globals).  Returns the list of (A_k, Pro_k) level views."""
    ctx = context(); Ad = _csr(A, ctx); keep = []
    o = _amg_options(amg_options, keep)
    J = C.c_int(0)
    ctx.call("ssn_amg_setup", C.byref(Ad.st), _byref_or_null(o), C.byref(J))
    levels = []
    for k in range(1, J.value + 1):
        a, p = CSR(), CSR()
        ctx.call("ssn_amg_level", k, C.byref(a), C.byref(p))
        levels.append((DeviceCSR(ctx, a, owned=False), DeviceCSR(ctx, p, owned=False) if k > 1 else None))
    return levels


def amg_clear():
    context().call("ssn_amg_clear")


def MG_Vcycle(r, isnsp=0, k=1):
    """``e = MG_Vcycle(r[,isnsp[,k]])`` -- reference AMG/MG_Vcycle.m:2-46 (live hierarchy)."""
    torch = _torch(); ctx = context(); host = _is_host(r)
    rd = _dev(r); e = torch.zeros_like(rd)
    ctx.call("ssn_mg_vcycle", _ptr(rd), int(isnsp), int(k), _ptr(e))
    return _ret(e, host)


def MG_Wcycle(r, isnsp=0, k=1, e=None):
    """``e = MG_Wcycle(r[,isnsp[,k[,e]]])`` -- reference AMG/MG_Wcycle.m:2-47 (live hierarchy)."""
    torch = _torch(); ctx = context(); host = _is_host(r, e)
    rd = _dev(r)
    ed = torch.zeros_like(rd) if e is None else _dev(e).clone()
    ctx.call("ssn_mg_wcycle", _ptr(rd), int(isnsp), int(k), _ptr(ed))
    return _ret(ed, host)


def Class_AMG(A, b, amg_options=None, keep_hierarchy=False):
    """``[x,it,rel_res,rel_resk,rhok] = Class_AMG(A,b[,amg_options])`` -- AMG/Class_AMG.m:1-111."""
    torch = _torch(); ctx = context(); host = _is_host(b); Ad = _csr(A, ctx); keep = []
    o = _amg_options(amg_options, keep)
    bd = _dev(b, count=Ad.shape[0])
    x = torch.empty_like(bd)
    maxit = 50 if o is None or o.maxit < 0 else o.maxit
    relk = np.zeros(maxit + 2); rhok = np.zeros(maxit + 2)
    it = C.c_int(0); rel = C.c_double(0.0); hl = C.c_int(0)
    ctx.call("ssn_class_amg", C.byref(Ad.st), _ptr(bd), _byref_or_null(o), 1 if keep_hierarchy else 0, _ptr(x),
             C.byref(it), C.byref(rel), relk.ctypes.data_as(C.c_void_p), rhok.ctypes.data_as(C.c_void_p), C.byref(hl))
    return _ret(x, host), it.value, rel.value, relk[:hl.value].copy(), rhok[:hl.value].copy()


def _twogrid_call(symbol, A, b, amg_options, nargin2_defaults, empty_defaults):
    torch = _torch(); ctx = context(); host = _is_host(b); Ad = _csr(A, ctx); keep = []
    if amg_options is None:
        amg_options = nargin2_defaults
    opts = dict(amg_options)
    for k, v in empty_defaults:
        if _empty(opts.get(k)):
            opts[k] = v
    o = _amg_options(opts, keep)
    bd = _dev(b, count=Ad.shape[0])
    x = torch.empty_like(bd)
    relk = np.zeros(o.maxit + 2); rhok = np.zeros(o.maxit + 2)
    it = C.c_int(0); rel = C.c_double(0.0); hl = C.c_int(0)
    ctx.call(symbol, C.byref(Ad.st), _ptr(bd), C.byref(o), _ptr(x), C.byref(it), C.byref(rel),
             relk.ctypes.data_as(C.c_void_p), rhok.ctypes.data_as(C.c_void_p), C.byref(hl))
    return _ret(x, host), it.value, rel.value, relk[:hl.value].copy(), rhok[:hl.value].copy()


def twogrid_bigph(A, b, amg_options=None):
    """``[x,it,rel_res,rel_resk,rhok] = twogrid_bigph(A,b[,amg_options])`` -- AMG/twogrid_bigph.m:1-116.
    ``nargin == 2`` defaults of :14-18 and the ``isempty`` defaults of :19-23 are applied here."""
    return _twogrid_call("ssn_twogrid_bigph", A, b, amg_options,
                         {"retol": 1e-12, "maxit": 20, "fnode": 0, "smoth": 10, "isnsp": 1, "guess": None},
                         (("retol", 0.0), ("maxit", 50), ("smoth", 3), ("isnsp", 0)))


def twogrid(A, b, amg_options=None):
    """``[x,it,rel_res,rel_resk,rhok] = twogrid(A,b[,amg_options])`` -- AMG/twogrid.m:1-150 (``nargin == 2``
    defaults of :16-21, ``isempty`` defaults of :22-34)."""
    return _twogrid_call("ssn_twogrid", A, b, amg_options,
                         {"retol": 1e-12, "bigph": 0, "maxit": 20, "smoth": 10, "isnsp": 1, "guess": None},
                         (("retol", 0.0), ("bigph", 0), ("maxit", 50), ("smoth", 3), ("isnsp", 0), ("fnode", 0)))


# ------------------------------------------------------------------ L2: Krylov

def PCG(H, e, pcg_options=None):
    """``[d,it,res,resk] = PCG(H,e[,pcg_options])`` -- reference PCG.m:1-105."""
    torch = _torch(); ctx = context(); host = _is_host(e); Hd = _csr(H, ctx); keep = []
    o = _pcg_options(pcg_options, keep)
    ed = _dev(e, count=Hd.shape[0]); d = torch.empty_like(ed)
    maxit = 10000 if o is None or o.maxit < 0 else o.maxit
    resk = np.zeros(max(maxit, 1))
    it = C.c_int(0); res = C.c_double(0.0)
    ctx.call("ssn_pcg", C.byref(Hd.st), _ptr(ed), _byref_or_null(o), _ptr(d), C.byref(it), C.byref(res),
             resk.ctypes.data_as(C.c_void_p))
    return _ret(d, host), it.value, res.value, resk[:maxit]


# ------------------------------------------------------------------ L3: dispatch

def components(A):
    """``[blocks,sizes,p,r] = components(A)`` -- reference components.m:1-64.  Frozen ordering:
    components by ascending smallest member, members ascending.  ``blocks`` 1-based labels,
    ``p`` 0-based node indices, ``r`` 0-based boundaries."""
    torch = _torch(); ctx = context(); Ad = _csr(A, ctx)
    if Ad.shape[0] != Ad.shape[1]:
        raise SsnError(-6, "Adjacency matrix must be square")
    N = Ad.shape[0]
    mk = lambda k: torch.empty(k, dtype=torch.int32, device="cuda")
    blocks, sizes, p, r = mk(N), mk(N), mk(N), mk(N + 1)
    nc = C.c_int(0)
    ctx.call("ssn_components", C.byref(Ad.st), _ptr(blocks), _ptr(sizes), _ptr(p), _ptr(r), C.byref(nc))
    k = nc.value
    return (blocks.cpu().numpy().astype(np.int64), sizes[:k].cpu().numpy().astype(np.int64),
            p.cpu().numpy().astype(np.int64), r[:k + 1].cpu().numpy().astype(np.int64))


def _prob_data(prob_data, ctx, keep, pot=False):
    torch = _torch()
    pd = ProbData()
    p, q = _dev(prob_data["p"]), _dev(prob_data["q"]); keep += [p, q]
    m, n = p.numel(), q.numel()
    pd.bk1, pd.tk, pd.m, pd.n = float(prob_data["bk1"]), float(prob_data["tk"]), m, n
    pd.p_dev, pd.q_dev = p.data_ptr(), q.data_ptr()
    T = prob_data.get("T")
    if T is not None:
        t = T.diagonal() if hasattr(T, "diagonal") and getattr(T, "ndim", 1) == 2 else T
        td = _dev(np.asarray(t) if not isinstance(t, torch.Tensor) else t, count=n + m); keep.append(td)
        pd.t_dev = td.data_ptr()
    H0 = _csr(prob_data["H0"], ctx); keep.append(H0)
    pd.H0 = C.pointer(H0.st)
    z = _dev(prob_data["z"], count=n + m + (1 if pot else 0)); keep.append(z)
    pd.z_dev = z.data_ptr()
    if pot:
        s = _dev(prob_data["s"], torch.uint8, count=m * n); phi = _dev(prob_data["phi"], count=m * n); keep += [s, phi]
        pd.s_dev, pd.phi_dev = s.data_ptr(), phi.data_ptr()
    return pd, m, n


def _solve(entry, prob_data, options, conv, pot=False, extra=()):
    torch = _torch(); ctx = context(); keep = []
    host = _is_host(prob_data["z"])
    pd, m, n = _prob_data(prob_data, ctx, keep, pot)
    o = conv(options, keep)
    zeta = torch.empty(n + m + (1 if pot else 0), dtype=torch.float64, device="cuda")
    it = C.c_int(0); res = C.c_double(0.0); info = (C.c_int * 2)()
    ctx.call(entry, C.byref(pd), _byref_or_null(o), *extra, _ptr(zeta), C.byref(it), C.byref(res), info)
    return _ret(zeta, host), it.value, res.value, np.array([info[0], info[1]])


def Hybrid_AMG(prob_data, amg_options):
    """``[zeta,itamg,resamg,info] = Hybrid_AMG(prob_data,amg_options)`` -- Hybrid_AMG.m:1-114."""
    return _solve("ssn_hybrid_amg", prob_data, amg_options, _amg_options)


def Hybrid_twogrid(prob_data, amg_options):
    """``[zeta,itamg,resamg,info] = Hybrid_twogrid(prob_data,amg_options)`` -- Hybrid_twogrid.m:1-90
    (``inner_solver = 5``)."""
    return _solve("ssn_hybrid_twogrid", prob_data, amg_options, _amg_options)


def ssn_step_class1(wk, lk, wlk, p, q, bk1, tk, gama=np.inf, inner_solver=4, amg_options=None, host_call=False):
    """One semismooth-Newton step of Class1/APD_SsN_Class1.m:137-212 at a fixed APD state as ONE library call
    (``ssn_ssn_step_class1``; ``host_call``: ``ssn_ssn_step_class1_host`` on host arrays, copies inside the call).
    Returns ``(lk_new, Fk_new, info)``."""
    torch = _torch(); ctx = context()
    keep = []
    o = _amg_options(amg_options, keep)
    info = (C.c_double * 12)()
    if host_call:
        arr = lambda v: v if (isinstance(v, torch.Tensor) and not v.is_cuda and v.dtype == torch.float64 and v.is_contiguous()) else \
            torch.from_numpy(np.ascontiguousarray(np.asarray(v.cpu() if hasattr(v, "cpu") else v, dtype=np.float64)))
        wh, lh, wlh, ph, qh = (arr(v) for v in (wk, lk, wlk, p, q))
        m, n = ph.numel(), qh.numel()
        gs = float(gama) if np.isscalar(gama) else float("inf")
        gh = None if np.isscalar(gama) else arr(gama)
        lk_new = torch.empty(n + m, dtype=torch.float64).pin_memory(); Fk_new = torch.empty(n + m, dtype=torch.float64).pin_memory()
        ctx.call("ssn_ssn_step_class1_host", wh.data_ptr(), lh.data_ptr(), wlh.data_ptr(), ph.data_ptr(), qh.data_ptr(), m, n, float(bk1),
                 float(tk), gh.data_ptr() if gh is not None else None, gs, int(inner_solver), _byref_or_null(o), lk_new.data_ptr(),
                 Fk_new.data_ptr(), C.cast(info, C.c_void_p))
    else:
        pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
        wd, ld, wld = _dev(wk, count=m * n), _dev(lk, count=n + m), _dev(wlk, count=n + m)
        gvec, gs = _gama_args(gama, m, n)
        lk_new = torch.empty(n + m, dtype=torch.float64, device="cuda"); Fk_new = torch.empty_like(lk_new)
        ctx.call("ssn_ssn_step_class1", _ptr(wd), _ptr(ld), _ptr(wld), _ptr(pd), _ptr(qd), m, n, float(bk1), float(tk), _ptr(gvec), gs,
                 int(inner_solver), _byref_or_null(o), _ptr(lk_new), _ptr(Fk_new), C.cast(info, C.c_void_p))
    v = list(info)
    return lk_new, Fk_new, {"E": int(v[0]), "nnzH": int(v[1]), "info": [int(v[2]), int(v[3])], "itamg": int(v[4]), "resamg": v[5], "ll": int(v[6]),
                            "ls_passes": int(v[7]), "Fk_old_norm": v[8], "Fk_new_norm": v[9], "ms_plan": v[10], "ms_amg": v[11], "ms_asat": None}


def warmup_class2(c, r, l, p, q, mu, phi, res=None, maxit=None):
    """``[uk,lk] = warmup_class2(c,r,l,p,q,mu,phi,res,maxit)`` -- reference Class2/warmup_class2.m:2-108 (A-ADMM warm start
    of partial OT), device resident (``ssn_warmup_class2``: two fused plan-wide kernels per iteration).  ``nargin`` rules of
    :3-18 as ``warmup_class1``; the residual test itself is commented out in the reference, so ``maxit`` iterations run."""
    torch = _torch(); ctx = context(); host = _is_host(c, r, l, p, q)
    if res is None:
        res = 1e-1
    if maxit is None:
        maxit = np.inf
    elif res == 0 and maxit == np.inf:
        raise ValueError("res = 0 and maxit = inf")
    if maxit == np.inf:
        maxit = 500
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
    cd = _dev(c, count=m * n); phid = _dev(phi, count=m * n)
    b = torch.cat([_dev(r), _dev(l), torch.tensor([float(mu)], dtype=torch.float64, device="cuda")])
    uk = torch.empty(m * n + n + m, dtype=torch.float64, device="cuda"); lk = torch.empty(n + m + 1, dtype=torch.float64, device="cuda")
    ctx.call("ssn_warmup_class2", _ptr(cd), _ptr(b), _ptr(pd), _ptr(qd), m, n, _ptr(phid), int(maxit), _ptr(uk), _ptr(lk))
    return _ret(uk, host), _ret(lk, host)


def apd_begin_pot(c, uk, vk, p, q, phi, b, lk, ak, bk, bk1):
    """``wk, huk, wlk`` of Class2/APD_SsN_Class2.m:121-122 in one pass over the x block."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel(); L = m * n + n + m
    cd, phid = _dev(c, count=m * n), _dev(phi, count=m * n)
    ud, vd, bd, ld = _dev(uk, count=L), _dev(vk, count=L), _dev(b, count=n + m + 1), _dev(lk, count=n + m + 1)
    wk = torch.empty(L, dtype=torch.float64, device="cuda"); huk = torch.empty(n + m + 1, dtype=torch.float64, device="cuda")
    wlk = torch.empty_like(huk)
    ctx.call("ssn_apd_begin_pot", _ptr(cd), _ptr(ud), _ptr(vd), _ptr(pd), _ptr(qd), m, n, _ptr(phid), _ptr(bd), _ptr(ld), float(ak), float(bk),
             float(bk1), _ptr(wk), _ptr(huk), _ptr(wlk))
    return wk, huk, wlk


def apd_end_pot(c, wk, uk, lk, p, q, phi, b, tk, ak):
    """``uk1, vk1, huk1`` and ``{c'xk1, KKT_xk^2, KKT_yk^2, KKT_zk^2, KKT_lk^2}`` of Class2/APD_SsN_Class2.m:231-238 in one pass."""
    torch = _torch(); ctx = context()
    pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel(); L = m * n + n + m
    cd, phid = _dev(c, count=m * n), _dev(phi, count=m * n)
    wd, ud, bd, ld = _dev(wk, count=L), _dev(uk, count=L), _dev(b, count=n + m + 1), _dev(lk, count=n + m + 1)
    uk1 = torch.empty(L, dtype=torch.float64, device="cuda"); vk1 = torch.empty_like(uk1)
    huk1 = torch.empty(n + m + 1, dtype=torch.float64, device="cuda")
    scal = (C.c_double * 5)()
    ctx.call("ssn_apd_end_pot", _ptr(cd), _ptr(wd), _ptr(ud), _ptr(ld), _ptr(pd), _ptr(qd), m, n, _ptr(phid), _ptr(bd), float(tk), float(ak),
             _ptr(uk1), _ptr(vk1), _ptr(huk1), C.cast(scal, C.c_void_p))
    return uk1, vk1, huk1, list(scal)


def APD_SsN_Class2(c, r, l, p, q, mu, phi, inner_solver=4, maxit=100, KKT_Tol=1e-6, warm_maxit=100, max_outer=None,
                   max_seconds=None, verbose=False, amg_options=None, pcg_options=None, host_call=False):
    """The reference's Class 2 script (Class2/APD_SsN_Class2.m:25-285 + Class2/warmup_class2.m) as ONE call into the library
    (``ssn_apd_ssn_class2``; ``host_call=True``: ``ssn_apd_ssn_class2_host`` on NumPy arrays).  Returns the dictionary of
    ``driver.APD_SsN_Class2``: ``KKT`` is the list of ``(KKT_xk, KKT_yk, KKT_zk, KKT_lk)`` per outer iteration."""
    from ._lib import ApdOptions, ApdResult
    torch = _torch(); ctx = context()
    keep = []
    ao = _amg_options(amg_options, keep); po = _pcg_options(pcg_options, keep)
    o = ApdOptions(inner_solver=int(inner_solver), maxit=int(maxit), KKT_Tol=float(KKT_Tol), warm_maxit=int(warm_maxit),
                   max_outer=int(max_outer or 0), max_seconds=float(max_seconds or 0.0), verbose=1 if verbose else 0,
                   amg=C.pointer(ao) if ao is not None else None, pcg=C.pointer(po) if po is not None else None)
    res = ApdResult()
    fx = np.zeros(int(maxit) + 1); kk = np.zeros((int(maxit) + 1, 4)); its = np.zeros(int(maxit), dtype=np.int32)
    cap = 64 * int(maxit)
    steps = np.zeros((cap, 7))
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    if host_call:
        f = lambda a: np.ascontiguousarray(np.asarray(a.cpu() if hasattr(a, "cpu") else a, dtype=np.float64).reshape(-1))
        ch, rh, lh, ph, qh, phih = f(c), f(r), f(l), f(p), f(q), f(phi); m, n = ph.size, qh.size
        uk = np.empty(m * n + n + m); lk = np.empty(m + n + 1)
        ctx.call("ssn_apd_ssn_class2_host", hp(ch), hp(rh), hp(lh), hp(ph), hp(qh), m, n, float(mu), hp(phih), C.byref(o), hp(uk), hp(lk),
                 C.byref(res), hp(fx), hp(kk), hp(its), hp(steps), cap)
    else:
        pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
        cd = _dev(c, count=m * n); rd = _dev(r, count=n); ld = _dev(l, count=m); phid = _dev(phi, count=m * n)
        uk = torch.empty(m * n + n + m, dtype=torch.float64, device="cuda"); lk = torch.empty(n + m + 1, dtype=torch.float64, device="cuda")
        ctx.call("ssn_apd_ssn_class2", _ptr(cd), _ptr(rd), _ptr(ld), _ptr(pd), _ptr(qd), m, n, float(mu), _ptr(phid), C.byref(o), _ptr(uk),
                 _ptr(lk), C.byref(res), hp(fx), hp(kk), hp(its), hp(steps), cap)
    L = res.hist_len; ns = min(int(res.steps_len), cap)
    stats = {"ssn_its": its[:res.outer_its].tolist(), "ls_trials": res.ls_trials, "converged": bool(res.converged),
             "amg_calls": res.amg_calls, "warmup_s": res.warmup_s, "solve_s": res.solve_s, "asat_s": res.asat_s, "plan_s": res.plan_s,
             "steps": [(int(a), int(b_), int(e), int(i0), int(it), int(ll), nf) for a, b_, e, i0, it, ll, nf in steps[:ns].tolist()]}
    return {"uk": uk, "xk": uk[:m * n], "lk": lk, "fxk": fx[:L].tolist(), "KKT": [tuple(row) for row in kk[:L].tolist()],
            "outer_its": res.outer_its, "rel_kkt": res.rel_kkt, "stats": stats, "seconds": res.loop_s, "warmup_seconds": res.warmup_s}


def ssn_step_class2(wk, lk, wlk, p, q, bk1, tk, phi, inner_solver=4, amg_options=None, pcg_options=None, host_call=False):
    """One semismooth-Newton step of Class2/APD_SsN_Class2.m:137-217 at a fixed APD state as ONE library call
    (``ssn_ssn_step_class2``; ``host_call``: ``ssn_ssn_step_class2_host`` on host arrays, copies inside the call).
    Returns ``(lk_new, Fk_new, info)``."""
    torch = _torch(); ctx = context()
    keep = []
    o = _amg_options(amg_options, keep); po = _pcg_options(pcg_options, keep)
    info = (C.c_double * 12)()
    if host_call:
        arr = lambda v: v if (isinstance(v, torch.Tensor) and not v.is_cuda and v.dtype == torch.float64 and v.is_contiguous()) else \
            torch.from_numpy(np.ascontiguousarray(np.asarray(v.cpu() if hasattr(v, "cpu") else v, dtype=np.float64)))
        wh, lh, wlh, ph, qh, phih = (arr(v) for v in (wk, lk, wlk, p, q, phi))
        m, n = ph.numel(), qh.numel()
        lk_new = torch.empty(n + m + 1, dtype=torch.float64).pin_memory(); Fk_new = torch.empty(n + m + 1, dtype=torch.float64).pin_memory()
        ctx.call("ssn_ssn_step_class2_host", wh.data_ptr(), lh.data_ptr(), wlh.data_ptr(), ph.data_ptr(), qh.data_ptr(), m, n, float(bk1),
                 float(tk), phih.data_ptr(), int(inner_solver), _byref_or_null(o), _byref_or_null(po), lk_new.data_ptr(), Fk_new.data_ptr(),
                 C.cast(info, C.c_void_p))
    else:
        pd, qd = _dev(p), _dev(q); m, n = pd.numel(), qd.numel()
        wd, ld, wld = _dev(wk, count=m * n + n + m), _dev(lk, count=n + m + 1), _dev(wlk, count=n + m + 1)
        phid = _dev(phi, count=m * n)
        lk_new = torch.empty(n + m + 1, dtype=torch.float64, device="cuda"); Fk_new = torch.empty_like(lk_new)
        ctx.call("ssn_ssn_step_class2", _ptr(wd), _ptr(ld), _ptr(wld), _ptr(pd), _ptr(qd), m, n, float(bk1), float(tk), _ptr(phid),
                 int(inner_solver), _byref_or_null(o), _byref_or_null(po), _ptr(lk_new), _ptr(Fk_new), C.cast(info, C.c_void_p))
    v = list(info)
    return lk_new, Fk_new, {"E": int(v[0]), "nnzH": int(v[1]), "info": [int(v[2]), int(v[3])], "itamg": int(v[4]), "resamg": v[5], "ll": int(v[6]),
                            "ls_passes": int(v[7]), "Fk_old_norm": v[8], "Fk_new_norm": v[9], "ms_plan": v[10], "ms_amg": v[11], "ms_asat": None}


def aug_PCG(prob_data, pcg_options):
    """``[zeta,itpcg,respcg,info] = aug_PCG(prob_data,pcg_options)`` -- aug_PCG.m:1-38."""
    return _solve("ssn_aug_pcg", prob_data, pcg_options, _pcg_options)


def AMG4POT(prob_data, amg_options, str_="amg"):
    """``[zeta,it,res,info] = AMG4POT(prob_data,amg_options,str)`` -- Class2/AMG4POT.m:1-56 (``str = 'amg'``: the two
    solves through Hybrid_AMG; anything else, e.g. ``'twogrid'``: through Hybrid_twogrid, :45-51)."""
    if str_ == "amg":
        return _solve("ssn_amg4pot", prob_data, amg_options, _amg_options, pot=True)
    return _solve("ssn_amg4pot_str", prob_data, amg_options, _amg_options, pot=True, extra=(1,))


def PCG4POT(prob_data, pcg_options):
    """``[zeta,it,res,info] = PCG4POT(prob_data,pcg_options)`` -- Class2/PCG4POT.m:1-40."""
    return _solve("ssn_pcg4pot", prob_data, pcg_options, _pcg_options, pot=True)


def rescaled_system(prob_data):
    """``Ae, f`` of Hybrid_AMG.m:17-24 (exported for parity tests)."""
    torch = _torch(); ctx = context(); keep = []
    pd, m, n = _prob_data(prob_data, ctx, keep)
    f = torch.empty(n + m, dtype=torch.float64, device="cuda"); st = CSR()
    ctx.call("ssn_rescaled_system", C.byref(pd), C.byref(st), _ptr(f))
    return DeviceCSR(ctx, st), f.cpu().numpy()


def jk_system(prob_data):
    """``Jk = bk1*speye(m+n) + (T+H0)/tk`` of Class1/APD_SsN_Class1.m:147,151, assembled on the device
    (``prob_data`` as for ``Hybrid_AMG``: ``bk1, tk, p, q, H0, z`` and optionally ``T``; only ``bk1, tk, T, H0`` are read)."""
    ctx = context(); keep = []
    pd, m, n = _prob_data(prob_data, ctx, keep)
    st = CSR()
    ctx.call("ssn_jk_system", C.byref(pd), C.byref(st))
    return DeviceCSR(ctx, st)


# ------------------------------------------------------------------ sparse utilities

def spmv(A, x):
    torch = _torch(); ctx = context(); host = _is_host(x); Ad = _csr(A, ctx)
    xd = _dev(x, count=Ad.shape[1]); y = torch.empty(Ad.shape[0], dtype=torch.float64, device="cuda")
    ctx.call("ssn_spmv", C.byref(Ad.st), _ptr(xd), _ptr(y))
    return _ret(y, host)


def spgemm(A, B):
    ctx = context(); Ad, Bd = _csr(A, ctx), _csr(B, ctx); st = CSR()
    ctx.call("ssn_spgemm", C.byref(Ad.st), C.byref(Bd.st), C.byref(st))
    return DeviceCSR(ctx, st)


def transpose(A):
    ctx = context(); Ad = _csr(A, ctx); st = CSR()
    ctx.call("ssn_transpose", C.byref(Ad.st), C.byref(st))
    return DeviceCSR(ctx, st)


def launch_count():
    return context().launches()


def kernel_timer(enable=True):
    """Switch on (and reset) / off the CUDA-event timer around the plan-wide kernel launches."""
    ctx = context(); ctx.call("ssn_kernel_timer", 1 if enable else 0)


def kernel_timer_read():
    """``(total_ms, launches)`` accumulated since ``kernel_timer(True)``."""
    ctx = context(); ms = C.c_double(0.0); cnt = C.c_int64(0)
    ctx.call("ssn_kernel_timer_read", C.byref(ms), C.byref(cnt))
    return ms.value, cnt.value


def set_persistent(enable=True):
    """Class_AMG's solve loop as one persistent cooperative kernel (default) or kernel by kernel."""
    ctx = context(); ctx.call("ssn_set_persistent", 1 if enable else 0)


def set_fused_setup(enable=True):
    """Small AMG levels coarsened by one kernel (default) or kernel by kernel -- ``ssn_set_fused_setup``."""
    context().call("ssn_set_fused_setup", 1 if enable else 0)


def set_cluster_solve(enable=True):
    """Class_AMG's solve loop inside one thread-block cluster (default) or grid-wide -- ``ssn_set_cluster_solve``."""
    context().call("ssn_set_cluster_solve", int(enable) if not isinstance(enable, bool) else (2 if enable else 0))


def set_spgemm_slab_limit(limit=1 << 30):
    """Intermediate-size limit above which a sparse product is formed in slabs of rows (``ssn_set_spgemm_slab_limit``)."""
    context().call("ssn_set_spgemm_slab_limit", int(limit))


def set_device_setup(enable=True):
    """PCG's SSOR / IC(0) factors and dependency levels built on the device (``True``) or on the host (default)."""
    ctx = context(); ctx.call("ssn_set_device_setup", 1 if enable else 0)


def set_dense_tail(enable=True, max_n=0):
    """Collapse the tail of small AMG levels into dense cycle operators (default) or walk it step by step."""
    ctx = context(); ctx.call("ssn_set_dense_tail", 1 if enable else 0, int(max_n))


def profile(enable=True):
    """Switch the library's phase profiler on/off (development aid)."""
    ctx = context(); ctx.lib.ssn_profile_enable(ctx.h, 1 if enable else 0)


def profile_dump():
    ctx = context(); return ctx.lib.ssn_profile_dump(ctx.h).decode()
