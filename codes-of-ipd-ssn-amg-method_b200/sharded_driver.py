"""Row-sharded caller of the inner solve: the reference's Class 1 script (Class1/APD_SsN_Class1.m,
Class1/warmup_class1.m) with every plan-sized array held as a row slab per rank (SURVEY.md
section 8e; BASELINE.json configs 4 and 5 -- the 256 x 256 grid is 34 GB per plan vector and only
exists sharded).

Rank g owns rows ``[g*m/G, (g+1)*m/G)`` of c, xk, vk, wk, ... as column-major ``m_loc x n`` slabs.
Plan-sized vector algebra is slab-local; ``Ax`` needs the path's one O(n) exchange (an
``all_reduce`` of the column sums, the row sums are all-gathered), scalars (norms, ``c'x``) ride in
small ``all_reduce`` calls; the semismooth-Newton step is ``sharded.ShardedStep`` (replicated AMG).
The operators come from ``ops`` (the CUDA operators by default; the CPU tests pass an adapter over
the oracle and run two gloo ranks).

The A-ADMM warm start runs the fused stage kernels of the single-GPU warm start on the slab
(``ssn_warm_stage``: 17 plan-sized reads / writes per iteration instead of ~45) with the column sums
exchanged between the two stages; the operator-by-operator form of ``driver.warmup_class1_unfused``
(one line per reference line, sharded ``Ax``) is kept as the cross-check (``fused_warmup=False``).
"""
import math
import time

import numpy as np

from .sharded import ShardedStep, _CudaOps, row_range


class SlabAlgebra:
    """``Ax`` / ``Aty`` / norms on row slabs, with the collectives they need."""

    def __init__(self, rank, world, p, q, ops, dist, torch):
        self.rank, self.world, self.ops, self.dist, self.torch = rank, world, ops, dist, torch
        self.p, self.q = p, q
        self.m, self.n = p.numel(), q.numel()
        self.r0, self.r1 = row_range(rank, world, self.m)
        self.m_loc = self.r1 - self.r0
        self.p_loc = p[self.r0:self.r1].contiguous()
        self.counts = [row_range(g, world, self.m)[1] - row_range(g, world, self.m)[0] for g in range(world)]
        self.collectives = 0

    def lam_loc(self, y):
        return self.torch.cat([y[: self.n], y[self.n + self.r0: self.n + self.r1]])

    def gather_rows(self, rows_loc):
        if self.world == 1:
            return rows_loc
        torch = self.torch
        mx = max(self.counts)
        pad = torch.zeros(mx, dtype=rows_loc.dtype, device=rows_loc.device)
        pad[: self.m_loc] = rows_loc
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad)
        self.collectives += 1
        return torch.cat([o[:c] for o, c in zip(out, self.counts)])

    def finish_ax(self, ax_loc, extra=None):
        """[column partials (n) ; local row sums (m_loc)] -> the full Ax vector (n+m); ``extra`` scalars
        (a list of floats) are summed over the ranks in the same message."""
        torch = self.torch
        cols = ax_loc[: self.n]
        if extra is not None:
            cols = torch.cat([cols, torch.tensor(list(extra), dtype=ax_loc.dtype, device=ax_loc.device)])
        if self.world > 1:
            cols = cols.contiguous()
            self.dist.all_reduce(cols)
            self.collectives += 1
        full = torch.cat([cols[: self.n], self.gather_rows(ax_loc[self.n:])])
        return (full, [float(v) for v in cols[self.n:]]) if extra is not None else full

    def Ax(self, x_loc):
        return self.finish_ax(self.ops.Ax(x_loc, self.p_loc, self.q))

    def Aty(self, y):
        return self.ops.Aty(self.lam_loc(y), self.p_loc, self.q)

    def sums(self, *vals):
        """all_reduce(sum) of a few host scalars."""
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.p.device)
        self.dist.all_reduce(t)
        self.collectives += 1
        return [float(v) for v in t]

    def maxs(self, *vals):
        """all_reduce(max) of a few host scalars."""
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.p.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        self.collectives += 1
        return [float(v) for v in t]


def warmup_class1_sharded_fused(A, c, b, gama, maxit):
    """Class1/warmup_class1.m:43-95 on row slabs with the fused stage kernels (``ops.warm_stage``): two
    plan-wide kernels per iteration on the slab, the column sums of ``Ax(dd)`` and of ``Ax(vk1), Ax(xk1)``
    exchanged in between (2 all_reduce + 2 all_gather calls per iteration), ``invAAt`` and the update of
    the dual block replicated on the (n+m)-vectors."""
    torch = A.torch
    ops, p, q = A.ops, A.p, A.q
    xk = torch.zeros_like(c); vk = torch.zeros_like(c); wk = torch.zeros_like(c); pik = torch.zeros_like(c)
    lk2 = torch.zeros_like(c); dd = torch.empty_like(c)
    lk1 = torch.zeros_like(b); axk = torch.zeros_like(b)
    b_loc = A.lam_loc(b)
    muf = 0.0; gk = 1.0; bk = 1.0                                       # :27
    for _ in range(int(maxit)):
        ak = bk; bk1 = bk / (1 + ak)                                    # :59
        gk1 = (gk + muf * ak) / (1 + ak)
        etafk = (1 + ak) * gk + muf * ak
        sgk = 1 / bk1
        tt = sgk * ak ** 2; sg = 1 + etafk / tt                         # :69
        ax_loc = ops.warm_stage(0, xk, vk, wk, pik, lk2, dd, c, A.p_loc, q, b_loc, A.lam_loc(lk1), A.lam_loc(axk), None,
                                ak, bk, gk, gama)                       # :63-67, Ax(dd)
        y = ops.invAAt(A.finish_ax(ax_loc), p, q, sg)                   # :70
        av_loc, axk_loc = ops.warm_stage(1, xk, vk, wk, pik, lk2, dd, c, A.p_loc, q, b_loc, None, None, A.lam_loc(y),
                                         ak, bk, gk, gama)              # :70-75, Ax(vk1), Ax(xk1)
        av = A.finish_ax(av_loc); axk = A.finish_ax(axk_loc)
        lk1 = lk1 + (ak / bk) * (av - b)                                # :75
        gk = gk1; bk = bk1                                              # :77
    return xk, lk1


def warmup_class1_sharded(A, c, b, gama, maxit):
    """Class1/warmup_class1.m:43-95 on row slabs (``A``: SlabAlgebra, ``c``: cost slab, ``b = [r ; l]``),
    operator by operator."""
    torch = A.torch
    p, q = A.p, A.q
    inf_gama = np.isscalar(gama) and math.isinf(gama)
    prox = (lambda x: torch.clamp_min(x, 0.0)) if inf_gama else (
        (lambda x: torch.clamp(x, 0.0, float(gama))) if np.isscalar(gama) else (lambda x: torch.minimum(torch.clamp_min(x, 0.0), gama)))
    Atb = A.Aty(b)
    muf = 0.0; gk = 1.0; bk = 1.0
    xk = torch.zeros_like(c); vk = xk.clone(); wk = xk.clone(); pik = xk.clone()
    lk1_ = torch.zeros_like(b); lk2_ = xk.clone()                       # lk = [lk1_ ; lk2_]
    for _ in range(int(maxit)):
        ak = bk; bk1 = bk / (1 + ak)
        gk1 = (gk + muf * ak) / (1 + ak)
        etafk = (1 + ak) * gk + muf * ak
        sgk = 1 / bk1; etagk = (1 + ak) * bk
        wwk = (ak * pik + wk) / (1 + ak)
        wxk = (ak * gk * vk + (gk + muf * ak) * xk) / etafk
        h1 = lk1_ - (A.Ax(xk) - b) / bk                                 # :65
        h2 = lk2_ - (xk - wk) / bk - (ak / bk) * (pik - wk)
        cAw = -Atb - wk; cAlk = A.Aty(h1) + h2                          # :66
        dd = etafk * wxk - ak ** 2 * (c + cAlk + sgk * cAw)             # :67
        del h2, cAw, cAlk, wxk
        tt = sgk * ak ** 2; sg = 1 + etafk / tt
        xk1 = (dd - A.Aty(A.ops.invAAt(A.Ax(dd), p, q, sg))) / (etafk + tt)   # :70
        del dd
        vk1 = xk1 + (xk1 - xk) / ak
        Av = A.Ax(vk1) - b
        blk2 = lk2_ + (ak / bk) * (vk1 - pik)                           # :72
        wk1 = prox(wwk - (ak ** 2 / etagk) * (-blk2))                   # :73
        del blk2, wwk
        pik1 = wk1 + (wk1 - wk) / ak
        lk1_ = lk1_ + (ak / bk) * Av                                    # :75
        lk2_ = lk2_ + (ak / bk) * (vk1 - pik1)
        gk = gk1; bk = bk1; xk = xk1; vk = vk1; wk = wk1; pik = pik1
    return xk, lk1_


def APD_SsN_Class1_sharded(c_loc, r, l, p, q, rank, world, gama=np.inf, maxit=100, KKT_Tol=1e-6, warm_maxit=100,
                           ops=None, dist=None, amg_options=None, max_outer=None, max_seconds=None, verbose=False, inner_solver=4,
                           fused_warmup=True, stop_on_amg_divergence=False):
    """APD outer loop + SsN inner loop of Class1/APD_SsN_Class1.m:32-275 on a row-sharded plan.
    ``c_loc``: this rank's slab of the cost (column-major ``m_loc x n``); ``r, l, p, q``: full vectors,
    replicated; ``inner_solver``: 4 = Hybrid_AMG (the reference's default), 5 = Hybrid_twogrid
    (Class1/APD_SsN_Class1.m:70,161,178).  Returns the same dictionary on every rank (``xk`` is the rank's slab).

    A divergent W-cycle solve (``Class_AMG`` leaving its loop on ``rho > 1``, AMG/Class_AMG.m:106) is recorded in
    ``stats["amg_diverged"]`` (first occurrence) and the iteration carries on with that direction, exactly like the
    reference (Class1/APD_SsN_Class1.m:161-212, where the line search decides) and like ``driver.APD_SsN_Class1``;
    ``stop_on_amg_divergence=True`` (off by default, a convenience for long multi-GPU runs) ends the solve there instead."""
    import torch
    if dist is None:
        import torch.distributed as dist
    ops = ops if ops is not None else _CudaOps()
    if amg_options is None:
        from .driver import CLASS1_AMG_OPTIONS as amg_options
    A = SlabAlgebra(rank, world, p, q, ops, dist, torch)
    n, m = A.n, A.m
    b = torch.cat([r, l])
    gam = float(gama)
    if not math.isinf(gam):
        raise NotImplementedError("the sharded outer loop covers gama = Inf (every shipped Class 1 configuration)")
    SsN_IT = 50; SsN_Tol1 = 1e-11                                       # :36
    sync = (lambda: torch.cuda.synchronize()) if torch.cuda.is_available() else (lambda: None)
    t_start = time.time()
    warm = warmup_class1_sharded_fused if (fused_warmup and hasattr(ops, "warm_stage")) else warmup_class1_sharded
    xk, lk = warm(A, c_loc, b, gama, warm_maxit)                         # :59
    sync(); t_warm = time.time() - t_start
    vk = xk.clone()

    def kkt(x_loc, lam):
        kl = float(torch.linalg.norm(A.Ax(x_loc) - b))
        px = ops.prox_residual(x_loc - c_loc, A.lam_loc(lam), A.p_loc, q, 1.0, gam, ("prox",))["prox"]   # prox(x-c-Aty(lam))
        d = x_loc - px
        kx2, cx = A.sums(float(d @ d), float(c_loc @ x_loc))
        return math.sqrt(kx2), kl, cx

    kx0, kl0, cx0 = kkt(xk, lk)
    fxk = [cx0]; KKT_xk = [kx0]; KKT_lk = [kl0]
    stats = {"ssn_its": [], "lin_its": [], "ls_trials": 0, "ls_passes": 0, "converged": False, "amg_calls": 0, "warmup_s": t_warm,
             "solve_ms": 0.0, "asat_ms": 0.0, "plan_ms": 0.0, "E": [], "steps": [], "amg_diverged": None}
    t_loop = time.time()
    bk = 1.0
    rr = [np.inf]
    k = 0
    for k in range(1, maxit + 1):                                       # :101
        resk = max(KKT_xk[k - 1], KKT_lk[k - 1])
        ak = math.sqrt(k ** 2 * bk)                                     # :113
        bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2               # :120
        SsN_Tol = max(bk1 / (k ** 2), SsN_Tol1)                         # :123
        wk, ax_loc = ops.apd_begin(c_loc, xk, vk, A.p_loc, q, ak, bk)    # :125 and Ax(xk) of :126, one pass over the slab
        axk = A.finish_ax(ax_loc)
        wlk = bk1 * (lk - 1 / bk * (axk - b)) - b                       # :126
        step = ShardedStep({"wk": wk, "lk": lk, "wlk": wlk, "p": p, "q": q, "bk1": bk1, "tk": tk, "gama": gam}, rank, world,
                           ops=ops, dist=dist, amg_options=amg_options, already_sharded=True, inner_solver=inner_solver)
        ssn_it = 0; lk_new = lk.clone(); stopped = False
        ev = step.residual(lk_new, True)                                # :129-130 (+ s for :140)
        Fk_new = bk1 * lk_new - ev[0] - wlk
        nF = float(torch.linalg.norm(Fk_new))
        Fk_res = nF
        its = []
        while nF > SsN_Tol:                                             # :137
            ssn_it += 1
            nFo = nF
            lk_new, Fk_new, info, ev = step.step(lk_new, pre=ev, want_s_new=True)    # :139-212
            stats["amg_calls"] += 1; stats["ls_trials"] += info["ll"] + 1; stats["ls_passes"] += info["ls_passes"]
            stats["solve_ms"] += info["ms_amg"]; stats["asat_ms"] += info["ms_asat"]; stats["plan_ms"] += info["ms_plan"]
            stats["E"].append(int(info["E"]))
            stats["steps"].append({"k": k, "ssn_it": ssn_it, "E": int(info["E"]), "nnzH": info["nnzH"], "amg_cycles": int(info["itamg"]),
                                   "amg_res": float(info["resamg"]), "ll": int(info["ll"]), "ms_plan": info["ms_plan"], "ms_asat": info["ms_asat"],
                                   "ms_amg": info["ms_amg"]})
            its.append(info["itamg"])
            nF = float(torch.linalg.norm(Fk_new))
            if not (info["resamg"] <= 1.0) or not math.isfinite(nF):
                # Class_AMG left its loop on rho > 1 (AMG/Class_AMG.m:106) with a residual above the initial one: the
                # W-cycle of the reference diverges on this system (the oracle reproduces it: the damped-Jacobi smoother
                # 0.5*D^-1 of a coarse Galerkin level has lambda_max(R*A) > 2).  The reference carries on with the
                # direction it got and so does this loop; the event is recorded.
                if stats["amg_diverged"] is None:
                    stats["amg_diverged"] = {"k": k, "ssn_it": ssn_it, "E": int(info["E"]), "amg_res": float(info["resamg"]), "Fk": nF}
                if stop_on_amg_divergence or not math.isfinite(nF):
                    stopped = True
                    break
            if verbose and rank == 0:
                print(f"   SsN: it={ssn_it:3d} |Fk|={nF:.2e} ll={info['ll']:3d} info={list(info['info'])} its={info['itamg']} "
                      f"res={info['resamg']:.2e} E={info['E']}", flush=True)
            if nF <= SsN_Tol:
                break
            if abs(nFo - nF) < SsN_Tol / 100:                           # :219
                break
            if ssn_it == SsN_IT:
                break
            if Fk_res / nF >= 2:
                Fk_res = nF
        A.collectives += step.collectives
        if stopped:
            stats["ssn_its"].append(ssn_it); stats["lin_its"].append(its)
            break
        lk1 = lk_new
        # :239-254 in one pass over the slab: xk1 = prox(zk), vk1, Ax(xk1), c'xk1 and the KKT residual of xk1
        xk1, vk1, ax1_loc, cx, kx2 = ops.apd_end(c_loc, wk, xk, A.lam_loc(lk1), A.p_loc, q, tk, ak, gam)
        axk1, (cx, kx2) = A.finish_ax(ax1_loc, extra=[cx, kx2])
        kl = float(torch.linalg.norm(axk1 - b)); kx = math.sqrt(kx2)
        rr = [kx / (1 + KKT_xk[0]), kl / (1 + KKT_lk[0])]
        if bk1 < 1e-8 and max(rr) > resk:                               # :245-249
            xk1 = xk; lk1 = lk; vk1 = xk; bk1 = float(ops.rand(1)[0])
            kx, kl, cx = kkt(xk1, lk1)
        bk = bk1; xk = xk1; lk = lk1; vk = vk1                          # :251
        del wk, step
        fxk.append(cx); KKT_lk.append(kl); KKT_xk.append(kx)
        stats["ssn_its"].append(ssn_it); stats["lin_its"].append(its)
        rr = [KKT_xk[k] / (1 + KKT_xk[0]), KKT_lk[k] / (1 + KKT_lk[0])]
        if verbose and rank == 0:
            print(f"APD: it={k:3d} KKT(xk)={rr[0]:.2e} KKT(lk)={rr[1]:.2e} fk={fxk[-1]:.8e} t={time.time() - t_loop:.2f}s", flush=True)
        if max(rr) <= KKT_Tol:                                          # :266
            stats["converged"] = True
            break
        if max_outer is not None and k >= max_outer:
            break
        if max_seconds is not None:
            # every rank must take the same branch: the slowest clock decides (all_reduce MAX)
            if A.maxs(time.time() - t_loop)[0] > max_seconds:
                break
    sync()
    stats["collectives"] = A.collectives
    return {"xk": xk, "lk": lk, "fxk": fxk, "KKT_xk": KKT_xk, "KKT_lk": KKT_lk, "outer_its": k, "rel_kkt": max(rr),
            "stats": stats, "seconds": time.time() - t_loop, "warmup_seconds": t_warm}


def grid_cost_slab(g, r0, r1, device="cuda"):
    """Rows ``[r0, r1)`` of the normalised squared-distance grid cost of ``problems.grid_cost`` as a
    column-major slab, generated on the device (the full 256 x 256 cost is 34 GB).  The arithmetic
    follows the host generator (``sq_i + sq_j - 2 x_i.y_j``, clamped at 0, divided by the maximum, which
    is attained between opposite corners); the matrix product is evaluated as two fused multiply-adds
    here and by BLAS there, so values agree to rounding, not to the bit."""
    import torch
    idx = torch.arange(g * g, device=device)
    pts = torch.stack([(torch.div(idx, g, rounding_mode="floor").double() + 0.5) / g, ((idx % g).double() + 0.5) / g], dim=1)
    sq = (pts ** 2).sum(1)
    rows = pts[r0:r1]
    C = sq[None, :] + sq[r0:r1][:, None] - 2.0 * (rows @ pts.T)          # [m_loc, n]
    C.clamp_(min=0.0)
    a, z = pts[0], pts[-1]
    cmax = float(sq[0] + sq[-1] - 2.0 * (a @ z))
    C /= cmax
    return C.t().contiguous().reshape(-1)                               # column-major m_loc x n: column j contiguous
