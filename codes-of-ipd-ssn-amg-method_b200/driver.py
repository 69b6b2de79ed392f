"""Device-resident caller of the inner solve: the reference's script bodies
(Class1/APD_SsN_Class1.m, Class1/warmup_class1.m) with every plan-sized array held as a torch
CUDA tensor (the analogue of running the MATLAB script on ``gpuArray``s) and every call on the
hot path going to the CUDA operators of this package.

The script is the CALLER of the hot path, kept here so that realistic
SsN states exist for the benchmark and so the full solve can be timed end to end.  The
semismooth-Newton residual, the active set and the line-search objective use the fused
``prox_residual`` kernel (one read of ``wk`` per evaluation) instead of the reference's
``Aty`` -> prox -> ``Ax`` chain; the arithmetic per entry is identical.
"""
import math
import os
import time

import numpy as np

from . import api

CLASS1_AMG_OPTIONS = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 1 / 4, "smoth": 5, "cycle": "w",
                      "isnsp": 1, "inter": 1, "guess": None}            # Class1/APD_SsN_Class1.m:87-88
CLASS1_PCG_OPTIONS = {"retol": 1e-11, "maxit": 10000, "precd": 2, "guess": None}   # :81


def _t(x, torch):
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=torch.float64).reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))).cuda()


def warmup_class1(c, r, l, p, q, gama, res=0.0, maxit=100):
    """A-ADMM warm start -- reference Class1/warmup_class1.m:18-96, on the device (fused kernels)."""
    import torch
    c, r, l, p, q = (_t(v, torch) for v in (c, r, l, p, q))
    return api.warmup_class1(c, r, l, p, q, gama, res, maxit)


def warmup_class1_unfused(c, r, l, p, q, gama, res=0.0, maxit=100):
    """The same warm start written operator by operator (Ax / Aty / invAAt calls and torch vector
    updates, one line per reference line) -- kept as the readable cross-check of the fused kernels."""
    import torch
    c, r, l, p, q = (_t(v, torch) for v in (c, r, l, p, q))
    m, n = l.numel(), r.numel()
    gam = None if (np.isscalar(gama) and math.isinf(gama)) else (gama if np.isscalar(gama) else _t(gama, torch))
    prox = (lambda x: torch.clamp_min(x, 0.0)) if gam is None else (lambda x: torch.minimum(torch.clamp_min(x, 0.0), gam) if not np.isscalar(gam) else torch.clamp(x, 0.0, gam))
    b = torch.cat([r, l]); Atb = api.Aty(b, p, q)
    muf = 0.0; gk = 1.0; bk = 1.0
    xk = torch.zeros(m * n, dtype=torch.float64, device="cuda"); vk = xk.clone(); wk = xk.clone(); pik = xk.clone()
    lk1_ = torch.zeros(m + n, dtype=torch.float64, device="cuda"); lk2_ = xk.clone()     # lk = [lk1_ ; lk2_]
    for _ in range(int(maxit)):                                         # warmup_class1.m:43-95
        ak = bk; bk1 = bk / (1 + ak)
        gk1 = (gk + muf * ak) / (1 + ak)
        etafk = (1 + ak) * gk + muf * ak
        sgk = 1 / bk1; etagk = (1 + ak) * bk
        wwk = (ak * pik + wk) / (1 + ak)
        wxk = (ak * gk * vk + (gk + muf * ak) * xk) / etafk
        h1 = lk1_ - (api.Ax(xk, p, q) - b) / bk                         # :65
        h2 = lk2_ - (xk - wk) / bk - (ak / bk) * (pik - wk)
        cAw = -Atb - wk; cAlk = api.Aty(h1, p, q) + h2                  # :66
        dd = etafk * wxk - ak ** 2 * (c + cAlk + sgk * cAw)             # :67
        del h2, cAw, cAlk, wxk
        tt = sgk * ak ** 2; sg = 1 + etafk / tt
        xk1 = (dd - api.Aty(api.invAAt(api.Ax(dd, p, q), p, q, sg), p, q)) / (etafk + tt)   # :70
        del dd
        vk1 = xk1 + (xk1 - xk) / ak
        Av = api.Ax(vk1, p, q) - b
        blk2 = lk2_ + (ak / bk) * (vk1 - pik)                           # :72
        wk1 = prox(wwk - (ak ** 2 / etagk) * (-blk2))                   # :73
        del blk2, wwk
        pik1 = wk1 + (wk1 - wk) / ak
        lk1_ = lk1_ + (ak / bk) * Av                                    # :75
        lk2_ = lk2_ + (ak / bk) * (vk1 - pik1)
        gk = gk1; bk = bk1; xk = xk1; vk = vk1; wk = wk1; pik = pik1
    return xk, lk1_


def APD_SsN_Class1(c, r, l, p, q, gama=np.inf, inner_solver=4, maxit=100, KKT_Tol=1e-6, warm_maxit=100,
                   on_ssn_step=None, verbose=False, max_outer=None, max_seconds=None, amg_options=None, pcg_options=None):
    """APD outer loop + SsN inner loop -- reference Class1/APD_SsN_Class1.m:32-275, on the device.

    ``on_ssn_step(state)`` is called right before every inner linear solve with a dict holding
    the tensors the solve reads (``wk, lk, wlk, bk1, tk, s, Fk, H0``).
    """
    import torch
    c, r, l, p, q = (_t(v, torch) for v in (c, r, l, p, q))
    m, n = l.numel(), r.numel()
    scalar_gama = np.isscalar(gama) or np.size(gama) == 1
    gam = float(gama) if scalar_gama else _t(gama, torch)
    b = torch.cat([r, l])
    bk = 1.0
    SsN_IT = 50; SsN_Tol1 = 1e-11; nu = 0.2; delta = 0.9; ll_max = 500   # :36
    amg_options = dict(amg_options or CLASS1_AMG_OPTIONS); pcg_options = dict(pcg_options or CLASS1_PCG_OPTIONS)
    t_start = time.time()
    xk, lk = warmup_class1(c, r, l, p, q, gama, 0.0, warm_maxit)         # :59
    torch.cuda.synchronize(); t_warm = time.time() - t_start
    vk = xk.clone()

    def kkt(x, lam):
        kl = float(torch.linalg.norm(api.Ax(x, p, q) - b))
        px = api.prox_residual(x - c, lam, p, q, 1.0, gam, want=("prox",))["prox"]           # prox(x-c-Aty(lam))
        kx = float(torch.linalg.norm(x - px))
        return kx, kl

    kx0, kl0 = kkt(xk, lk)
    fxk = [float(c @ xk)]; KKT_xk = [kx0]; KKT_lk = [kl0]
    stats = {"ssn_its": [], "lin_its": [], "ls_trials": 0, "converged": False, "amg_calls": 0, "warmup_s": t_warm,
             "solve_s": 0.0, "asat_s": 0.0, "plan_s": 0.0, "ls_passes": 0, "solve_calls": [],
             "steps": []}                 # (k, ssn_it, E, components, inner its, ll, |Fk_new|) per SsN step, as the oracle records them
    t_loop = time.time()
    rr = [np.inf]
    k = 0
    for k in range(1, maxit + 1):                                       # :101
        resk = max(KKT_xk[k - 1], KKT_lk[k - 1])
        ak = math.sqrt(k ** 2 * bk)                                     # :113
        bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2               # :120
        SsN_Tol = max(bk1 / (k ** 2), SsN_Tol1)                         # :123
        wk, axk = api.apd_begin(c, xk, vk, p, q, ak, bk)               # :125 and Ax(xk) of :126, one pass
        wlk = bk1 * (lk - 1 / bk * (axk - b)) - b                       # :126
        ssn_it = 0; lk_new = lk.clone()
        ev = api.prox_residual(wk, lk_new, p, q, tk, gam, want=("Axprox", "s"))   # :129-130 (+ s for :140)
        Fk_new = bk1 * lk_new - ev["Axprox"] - wlk
        nF = float(torch.linalg.norm(Fk_new))
        Fk_res = nF
        its = []
        while nF > SsN_Tol:                                             # :137
            ssn_it += 1; lk_old = lk_new
            Fk_old = Fk_new; s = ev["s"]; n2_old = ev["norm2"]          # z, s, prox of lk_old: already evaluated
            t0 = time.time()
            H0 = api.ASAt(s, p, q)                                      # :142
            torch.cuda.synchronize(); stats["asat_s"] += time.time() - t0
            if on_ssn_step is not None:
                on_ssn_step({"k": k, "ssn_it": ssn_it, "wk": wk, "lk": lk_old, "wlk": wlk, "bk1": bk1, "tk": tk,
                             "s": s, "Fk": Fk_old, "H0": H0, "E": ev["count"], "xk": xk, "vk": vk, "ak": ak, "bk": bk})
            t0 = time.time()
            prob_data = {"bk1": bk1, "tk": tk, "q": q, "p": p, "T": None, "H0": H0, "z": -Fk_old}
            if inner_solver == 2:                                       # :149-152, PCG on Jk = bk1*I + (T+H0)/tk
                po = dict(pcg_options)
                if po.get("precd") == 5:
                    po["nf"] = n
                Jk = api.jk_system(prob_data)                           # ssn_jk_system: assembled on the device
                zeta, itpcg, respcg, _ = api.PCG(Jk, -Fk_old, po)
                info = [0, 0]
            elif inner_solver == 3:
                zeta, itpcg, respcg, info = api.aug_PCG(prob_data, pcg_options)
            elif inner_solver == 4:
                zeta, itpcg, respcg, info = api.Hybrid_AMG(prob_data, amg_options)
                stats["amg_calls"] += 1
            elif inner_solver == 5:
                zeta, itpcg, respcg, info = api.Hybrid_twogrid(prob_data, amg_options)          # :178
                stats["amg_calls"] += 1
            else:
                raise ValueError("inner_solver must be 2 (PCG), 3 (aug_PCG), 4 (Hybrid_AMG) or 5 (Hybrid_twogrid)")
            torch.cuda.synchronize(); stats["solve_s"] += time.time() - t0
            stats["solve_calls"].append((int(ev["count"]), time.time() - t0, int(itpcg), int(info[0])))
            its.append(itpcg); E_step = int(ev["count"])
            t0 = time.time()
            f0 = bk1 / 2 * float(lk_old @ lk_old) - float(wlk @ lk_old)  # :182
            cFk_old = f0 + 0.5 * tk * n2_old
            ress = abs(float(Fk_old @ zeta))
            # :189-211, ll = 0 alone, then 8-128 backtracking steps per read of wk (adaptive, api.linesearch)
            lk_new, ll, _, _, passes = api.linesearch(wk, lk_old, zeta, wlk, p, q, tk, bk1, cFk_old, ress, gam, nu, delta, ll_max)
            stats["ls_trials"] += ll + 1; stats["ls_passes"] += passes
            ev = api.prox_residual(wk, lk_new, p, q, tk, gam, want=("Axprox", "s"))
            Fk_new = bk1 * lk_new - ev["Axprox"] - wlk                  # :212
            nFo = float(torch.linalg.norm(Fk_old)); nF = float(torch.linalg.norm(Fk_new))
            torch.cuda.synchronize(); stats["plan_s"] += time.time() - t0
            stats["steps"].append((k, ssn_it, E_step, int(info[0]), int(itpcg), int(ll), nF))
            if verbose:
                print(f"   SsN: it={ssn_it:3d} |Fk|={nF:.2e} ll={ll:3d} info={list(info)} its={itpcg} res={respcg:.2e} E={ev['count']}")
            if nF <= SsN_Tol:
                break
            if abs(nFo - nF) < SsN_Tol / 100:                           # :219
                break
            if ssn_it == SsN_IT:
                break
            if Fk_res / nF >= 2:
                Fk_res = nF
        lk1 = lk_new
        # :239-254 in one pass: xk1 = prox(zk), vk1, Ax(xk1), c'xk1 and the KKT residual of xk1
        xk1, vk1, axk1, cx, kx2 = api.apd_end(c, wk, xk, lk1, p, q, tk, ak, gam)
        kl = float(torch.linalg.norm(axk1 - b)); kx = math.sqrt(kx2)
        rr = [kx / (1 + KKT_xk[0]), kl / (1 + KKT_lk[0])]
        if bk1 < 1e-8 and max(rr) > resk:                               # :245-249
            xk1 = xk; lk1 = lk; vk1 = xk; bk1 = float(api.rand(1)[0])
            kx, kl = kkt(xk1, lk1); cx = float(c @ xk1)
        bk = bk1; xk = xk1; lk = lk1; vk = vk1                          # :251
        fxk.append(cx); KKT_lk.append(kl); KKT_xk.append(kx)
        stats["ssn_its"].append(ssn_it); stats["lin_its"].append(its)
        rr = [KKT_xk[k] / (1 + KKT_xk[0]), KKT_lk[k] / (1 + KKT_lk[0])]
        if verbose:
            print(f"APD: it={k:3d} KKT(xk)={rr[0]:.2e} KKT(lk)={rr[1]:.2e} fk={fxk[-1]:.8e} t={time.time() - t_loop:.2f}s")
        if max(rr) <= KKT_Tol:                                          # :266
            stats["converged"] = True
            break
        if max_outer is not None and k >= max_outer:
            break
        if max_seconds is not None and time.time() - t_loop > max_seconds:
            break
    torch.cuda.synchronize()
    return {"xk": xk, "lk": lk, "fxk": fxk, "KKT_xk": KKT_xk, "KKT_lk": KKT_lk, "outer_its": k, "rel_kkt": max(rr),
            "stats": stats, "seconds": time.time() - t_loop, "warmup_seconds": t_warm}


# ------------------------------------------------------------------ Class 2: partial optimal transport

CLASS2_AMG_OPTIONS = {"retol": 1e-11, "bigph": 1, "maxit": 40, "theta": 1 / 4, "smoth": 10, "cycle": "w",
                      "isnsp": 1, "inter": 1, "guess": None}            # Class2/APD_SsN_Class2.m:80-81


def warmup_class2(c, r, l, p, q, mu, phi, res=0.0, maxit=100):
    """A-ADMM warm start for partial OT -- reference Class2/warmup_class2.m:18-108, on the device (fused kernels,
    ``ssn_warmup_class2``)."""
    import torch
    c, r, l, p, q, phi = (_t(v, torch) for v in (c, r, l, p, q, phi))
    return api.warmup_class2(c, r, l, p, q, mu, phi, res, maxit)


def warmup_class2_unfused(c, r, l, p, q, mu, phi, res=0.0, maxit=100):
    """The same warm start operator by operator (Ax, Aty, invHHt and torch vector updates): the cross-check of the fused one."""
    import torch
    c, r, l, p, q, phi = (_t(v, torch) for v in (c, r, l, p, q, phi))
    m, n = l.numel(), r.numel(); N = m + n; mn = m * n
    f64 = dict(dtype=torch.float64, device="cuda")
    b = torch.cat([r, l, torch.tensor([float(mu)], **f64)])
    Hmul = lambda u: torch.cat([api.Ax(u[:mn], p, q) + u[mn:], (phi @ u[:mn]).reshape(1)])
    Htmul = lambda lam: torch.cat([api.Aty(lam[:N], p, q) + lam[N] * phi, lam[:N]])
    Htb = Htmul(b)                                                      # :22
    wc = torch.cat([c, torch.zeros(N, **f64)])
    muf = 0.0; gk = 1.0; bk = 1.0
    uk = torch.zeros(mn + N, **f64); vk = uk.clone(); wk = uk.clone(); pik = uk.clone()
    lkA = torch.zeros(N + 1, **f64); lkB = uk.clone()                   # lk = [lkA ; lkB], :26
    for _ in range(int(maxit)):                                         # :46-107
        ak = bk; bk1 = bk / (1 + ak)
        gk1 = (gk + muf * ak) / (1 + ak)
        etafk = (1 + ak) * gk + muf * ak
        sgk = 1 / bk1; etagk = (1 + ak) * bk
        wwk = (ak * pik + wk) / (1 + ak)
        wuk = (ak * gk * vk + (gk + muf * ak) * uk) / etafk
        hA = lkA - (Hmul(uk) - b) / bk                                  # :66
        hB = lkB - (uk - wk) / bk - (ak / bk) * (pik - wk)
        cAw = -Htb - wk
        cAlk = hB + Htmul(hA)                                           # :67-68
        dd = etafk * wuk - ak ** 2 * (wc + cAlk + sgk * cAw)            # :69
        del hB, cAw, cAlk, wuk
        tt = sgk * ak ** 2; sg = 1 + etafk / tt
        ff = api.invHHt(Hmul(dd), p, q, sg, phi)                        # :71-72
        uk1 = (dd - Htmul(ff)) / (etafk + tt)                           # :73
        del dd
        vk1 = uk1 + (uk1 - uk) / ak
        b0 = Hmul(vk1) - b                                              # :75
        blkB = lkB + (ak / bk) * (vk1 - pik)
        wk1 = torch.clamp_min(wwk - (ak ** 2 / etagk) * (-blkB), 0.0)   # :77
        del blkB, wwk
        pik1 = wk1 + (wk1 - wk) / ak
        lkA = lkA + (ak / bk) * b0                                      # :79
        lkB = lkB + (ak / bk) * (vk1 - pik1)
        gk = gk1; bk = bk1; uk = uk1; vk = vk1; wk = wk1; pik = pik1
    return uk, lkA


def APD_SsN_Class2(c, r, l, p, q, mu, phi, inner_solver=4, maxit=100, KKT_Tol=1e-6, warm_maxit=100,
                   on_ssn_step=None, verbose=False, max_outer=None, max_seconds=None, amg_options=None):
    """APD outer loop + SsN inner loop for partial OT -- reference Class2/APD_SsN_Class2.m:25-285 on the device (inner
    solver 3 = PCG4POT, 4 = AMG4POT, 5 = AMG4POT with 'twogrid').  u = [x (mn); y (n); z (m)], duals lk (n+m+1).  Without
    an ``on_ssn_step`` hook the whole script is ONE library call (``ssn_apd_ssn_class2``); with one, the Python loop below
    runs over the fused operators so that the hook can see every state."""
    import torch
    if on_ssn_step is None:
        return api.APD_SsN_Class2(c, r, l, p, q, mu, phi, inner_solver=inner_solver, maxit=maxit, KKT_Tol=KKT_Tol, warm_maxit=warm_maxit,
                                  max_outer=max_outer, max_seconds=max_seconds, verbose=verbose, amg_options=amg_options)
    return APD_SsN_Class2_loop(c, r, l, p, q, mu, phi, inner_solver, maxit, KKT_Tol, warm_maxit, on_ssn_step, verbose, max_outer,
                               max_seconds, amg_options)


def APD_SsN_Class2_loop(c, r, l, p, q, mu, phi, inner_solver=4, maxit=100, KKT_Tol=1e-6, warm_maxit=100,
                        on_ssn_step=None, verbose=False, max_outer=None, max_seconds=None, amg_options=None, fused_warmup=True):
    """The script as a Python loop over the library's operators (the form the one-call entry point replaced): kept for the
    ``on_ssn_step`` hook and as the cross-check of ``ssn_apd_ssn_class2``."""
    import torch
    c, r, l, p, q, phi = (_t(v, torch) for v in (c, r, l, p, q, phi))
    m, n = l.numel(), r.numel(); N = m + n; mn = m * n
    f64 = dict(dtype=torch.float64, device="cuda")
    b = torch.cat([r, l, torch.tensor([float(mu)], **f64)]); wc = torch.cat([c, torch.zeros(N, **f64)])
    bk = 1.0
    SsN_IT = 50; SsN_Tol1 = 1e-10; nu = 0.2; delta = 0.9; ll_max = 500   # :28
    amg_options = dict(amg_options or CLASS2_AMG_OPTIONS); pcg_options = dict(CLASS1_PCG_OPTIONS)
    Hmul = lambda u: torch.cat([api.Ax(u[:mn], p, q) + u[mn:], (phi @ u[:mn]).reshape(1)])
    Htmul = lambda lam: torch.cat([api.Aty(lam[:N], p, q) + lam[N] * phi, lam[:N]])
    nrm = lambda v: float(torch.linalg.norm(v))
    t_start = time.time()
    uk, lk = (warmup_class2 if fused_warmup else warmup_class2_unfused)(c, r, l, p, q, mu, phi, 0.0, warm_maxit)     # :50
    torch.cuda.synchronize(); t_warm = time.time() - t_start
    vk = uk.clone()

    def kkts(u, lam):
        x, y, z = u[:mn], u[mn:mn + n], u[mn + n:]
        return (nrm(x - torch.clamp_min(x - c - (api.Aty(lam[:N], p, q) + lam[N] * phi), 0.0)),
                nrm(y - torch.clamp_min(y - lam[:n], 0.0)), nrm(z - torch.clamp_min(z - lam[n:N], 0.0)),
                nrm(Hmul(u) - b))

    KKT = [kkts(uk, lk)]; fxk = [float(c @ uk[:mn])]
    stats = {"ssn_its": [], "lin_its": [], "ls_trials": 0, "converged": False, "amg_calls": 0, "warmup_s": t_warm, "steps": []}
    t_loop = time.time(); rr = [np.inf]; k = 0
    for k in range(1, maxit + 1):                                       # :95
        resk = max(KKT[k - 1])
        ak = math.sqrt(k ** 2 * bk)                                     # :116
        bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2
        SsN_Tol = max(bk1 / (k ** 2), SsN_Tol1)
        wk = -wc + bk * (uk + ak * vk) / ak ** 2                        # :121
        wlk = bk1 * (lk - 1 / bk * (Hmul(uk) - b)) - b                  # :122
        ssn_it = 0; lk_new = lk.clone()
        # zk = 1/tk*(wk - Htmul(lk)), prox, H*prox, ||prox||^2 and the active flags in ONE fused pass over wk and phi
        # (ssn_prox_residual_pot) instead of ~10 plan-sized element-wise passes                          :127-130
        ev = api.prox_residual_pot(wk, lk_new, p, q, tk, phi, want=("Hprox", "s", "t"))
        Fk_new = bk1 * lk_new - ev["Hprox"] - wlk                       # :130
        nF = nrm(Fk_new); Fk_res = nF
        its = []
        while nF > SsN_Tol:                                             # :136
            ssn_it += 1; lk_old = lk_new; Fk_old = Fk_new               # ev is already the evaluation at lk_old
            s = ev["s"]; t = ev["t"]                                    # :139
            H0 = api.ASAt(s, p, q)                                      # :146
            prob_data = {"bk1": bk1, "tk": tk, "q": q, "p": p, "s": s, "T": t, "H0": H0, "z": -Fk_old, "phi": phi}
            if on_ssn_step is not None:
                on_ssn_step(dict(prob_data, k=k, ssn_it=ssn_it))
            if inner_solver == 3:
                zeta, itpcg, respcg, info = api.PCG4POT(prob_data, pcg_options)        # :168
            elif inner_solver in (4, 5):
                zeta, itpcg, respcg, info = api.AMG4POT(prob_data, amg_options, "amg" if inner_solver == 4 else "twogrid")   # :171 / :182
                stats["amg_calls"] += 1
            else:
                raise ValueError("inner_solver must be 3 (PCG4POT), 4 (AMG4POT) or 5 (AMG4POT, 'twogrid')")
            its.append(itpcg)
            cFk_old = bk1 / 2 * float(lk_old @ lk_old) - float(wlk @ lk_old) + 0.5 * tk * ev["norm2"]   # :196-197
            ress = abs(float(Fk_old @ zeta))
            ll = 0
            while True:                                                 # :199-213, one fused pass per trial
                lk_new = lk_old + delta ** ll * zeta
                f0 = bk1 / 2 * float(lk_new @ lk_new) - float(wlk @ lk_new)
                ev = api.prox_residual_pot(wk, lk_new, p, q, tk, phi, want=("Hprox", "s", "t"))
                if not (f0 + 0.5 * tk * ev["norm2"] > cFk_old - nu * delta ** ll * ress) or ll == ll_max:
                    break
                ll += 1
            stats["ls_trials"] += ll + 1
            Fk_new = bk1 * lk_new - ev["Hprox"] - wlk                   # :217
            nFo = nrm(Fk_old); nF = nrm(Fk_new)
            stats["steps"].append((k, ssn_it, int(s.sum()), int(info[0]), int(itpcg), int(ll), nF))
            if verbose:
                print(f"   SsN: it={ssn_it:3d} |Fk|={nF:.2e} ll={ll:3d} info={list(info)} its={itpcg} res={respcg:.2e}")
            if nF <= SsN_Tol:
                break
            if abs(nFo - nF) < SsN_Tol:                                 # :224
                break
            if ssn_it == SsN_IT:
                break
            if Fk_res / nF >= 2:
                Fk_res = nF
        lk1 = lk_new
        uk1 = api.prox_residual_pot(wk, lk_new, p, q, tk, phi, want=("prox",))["prox"]   # prox(zk) at the accepted duals
        vk1 = uk1 + (uk1 - uk) / ak                                     # :244
        kk = kkts(uk1, lk1)
        rr = [kk[i] / (1 + KKT[0][i]) for i in range(4)]
        if bk1 < 1e-8 and max(rr) > resk:                               # :253-257
            uk1 = uk; lk1 = lk; vk1 = uk; bk1 = 10 * bk1
            kk = kkts(uk1, lk1)
        bk = bk1; uk = uk1; lk = lk1; vk = vk1                          # :259
        fxk.append(float(c @ uk[:mn])); KKT.append(kk)
        stats["ssn_its"].append(ssn_it); stats["lin_its"].append(its)
        rr = [KKT[k][i] / (1 + KKT[0][i]) for i in range(4)]
        if verbose:
            print(f"APD: it={k:3d} KKT(x,y,z,l)={['%.2e' % v for v in rr]} fk={fxk[-1]:.8e} t={time.time() - t_loop:.2f}s")
        if max(rr) <= KKT_Tol:                                          # :274
            stats["converged"] = True
            break
        if max_outer is not None and k >= max_outer:
            break
        if max_seconds is not None and time.time() - t_loop > max_seconds:
            break
    torch.cuda.synchronize()
    return {"uk": uk, "xk": uk[:mn], "lk": lk, "fxk": fxk, "KKT": KKT, "outer_its": k, "rel_kkt": max(rr), "stats": stats,
            "seconds": time.time() - t_loop, "warmup_seconds": t_warm}


def class2_trivial_state(P):
    """The APD state of outer iteration 1 of Class2/APD_SsN_Class2.m from the trivial start (``uk = vk = 0, lk = 0, bk = 1``;
    :116-122), on the device: what ``ssn_step_class2`` takes."""
    import torch
    f64 = dict(dtype=torch.float64, device="cuda")
    c, r, l, p, q, phi = (_t(P[k], torch) for k in ("c", "r", "l", "p", "q", "phi"))
    m, n = l.numel(), r.numel()
    b = torch.cat([r, l, torch.tensor([float(P["mu"])], **f64)])
    ak = 1.0; bk = 1.0
    bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2
    wk = -torch.cat([c, torch.zeros(m + n, **f64)])
    wlk = bk1 * (torch.zeros(m + n + 1, **f64) - 1 / bk * (0.0 - b)) - b
    return {"wk": wk, "lk": torch.zeros(m + n + 1, **f64), "wlk": wlk, "p": p, "q": q, "phi": phi, "tk": tk, "bk1": bk1, "m": m, "n": n, "k": 1}


def ssn_step_class2(state, amg_options=None, host_call=False):
    """One semismooth-Newton step of Class2/APD_SsN_Class2.m:137-217 at a fixed APD state as ONE library call
    (``ssn_ssn_step_class2``).  Returns ``(lk_new, Fk_new, info)``."""
    return api.ssn_step_class2(state["wk"], state["lk"], state["wlk"], state["p"], state["q"], state["bk1"], state["tk"], state["phi"],
                               inner_solver=4, amg_options=amg_options or CLASS2_AMG_OPTIONS, host_call=host_call)


def ssn_step_class2_ops(state, amg_options=None, max_ll=500):
    """The same step operator by operator with host-timed phases: fused residual + active flags (``ssn_prox_residual_pot``)
    -> ASAt -> AMG4POT -> Armijo line search (one fused pass per trial) -> new residual.  Returns ``(lk_new, Fk_new, info)``."""
    import torch
    wk, lk, wlk, p, q, phi = state["wk"], state["lk"], state["wlk"], state["p"], state["q"], state["phi"]
    bk1, tk = state["bk1"], state["tk"]
    nu, delta = 0.2, 0.9
    tm = {"plan": 0.0, "asat": 0.0, "amg": 0.0}

    def lap(key, t0):
        torch.cuda.synchronize(); tm[key] += (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    ev = api.prox_residual_pot(wk, lk, p, q, tk, phi, want=("Hprox", "s", "t"))      # :139-150
    Fk_old = bk1 * lk - ev["Hprox"] - wlk
    lap("plan", t0); t0 = time.perf_counter()
    H0 = api.ASAt(ev["s"], p, q)                                                     # :146
    lap("asat", t0); t0 = time.perf_counter()
    prob_data = {"bk1": bk1, "tk": tk, "q": q, "p": p, "s": ev["s"], "T": ev["t"], "H0": H0, "z": -Fk_old, "phi": phi}
    zeta, itamg, resamg, info = api.AMG4POT(prob_data, amg_options or CLASS2_AMG_OPTIONS, "amg")   # :171
    lap("amg", t0); t0 = time.perf_counter()
    cFk_old = bk1 / 2 * float(lk @ lk) - float(wlk @ lk) + 0.5 * tk * ev["norm2"]    # :196-197
    ress = abs(float(Fk_old @ zeta))
    ll = 0
    while True:                                                                      # :199-213
        lk_new = lk + delta ** ll * zeta
        f0 = bk1 / 2 * float(lk_new @ lk_new) - float(wlk @ lk_new)
        ev2 = api.prox_residual_pot(wk, lk_new, p, q, tk, phi, want=("Hprox",))
        if not (f0 + 0.5 * tk * ev2["norm2"] > cFk_old - nu * delta ** ll * ress) or ll == max_ll:
            break
        ll += 1
    Fk_new = bk1 * lk_new - ev2["Hprox"] - wlk                                       # :217
    lap("plan", t0)
    return lk_new, Fk_new, {"E": ev["count"], "ms_plan": tm["plan"], "ms_asat": tm["asat"], "ms_amg": tm["amg"], "itamg": itamg,
                            "resamg": resamg, "info": info, "ll": ll, "ls_passes": ll + 1, "nnzH": H0.nnz,
                            "Fk_old_norm": float(torch.linalg.norm(Fk_old)), "Fk_new_norm": float(torch.linalg.norm(Fk_new))}


def ssn_step(state, amg_options=None, max_ll=500):
    """One semismooth-Newton step of Class1/APD_SsN_Class1.m:137-212 at a fixed APD state:
    fused residual + active set -> ASAt -> Hybrid_AMG -> Armijo line search -> new residual.
    ``state`` holds device tensors ``wk, lk, wlk, p, q`` and scalars ``bk1, tk, gama``.
    Returns ``(lk_new, Fk_new, info)``."""
    import torch
    wk, lk, wlk, p, q = state["wk"], state["lk"], state["wlk"], state["p"], state["q"]
    bk1, tk, gam = state["bk1"], state["tk"], state.get("gama", float("inf"))
    nu, delta = 0.2, 0.9
    tm = {"plan": 0.0, "asat": 0.0, "amg": 0.0}

    def lap(key, t0):
        torch.cuda.synchronize(); tm[key] += (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    ev = api.prox_residual(wk, lk, p, q, tk, gam, want=("Axprox", "s"))          # :139-144
    Fk_old = bk1 * lk - ev["Axprox"] - wlk
    lap("plan", t0); t0 = time.perf_counter()
    H0 = api.ASAt(ev["s"], p, q)                                                 # :142
    lap("asat", t0); t0 = time.perf_counter()
    prob_data = {"bk1": bk1, "tk": tk, "q": q, "p": p, "T": None, "H0": H0, "z": -Fk_old}
    zeta, itamg, resamg, info = api.Hybrid_AMG(prob_data, amg_options or CLASS1_AMG_OPTIONS)   # :161
    lap("amg", t0); t0 = time.perf_counter()
    f0 = bk1 / 2 * float(lk @ lk) - float(wlk @ lk)                              # :182-184
    cFk_old = f0 + 0.5 * tk * ev["norm2"]
    ress = abs(float(Fk_old @ zeta))
    # :189-211, ll = 0 alone, then 8-32 backtracking steps per read of wk (adaptive, api.linesearch)
    lk_new, ll, _, _, passes = api.linesearch(wk, lk, zeta, wlk, p, q, tk, bk1, cFk_old, ress, gam, nu, delta, max_ll)
    ev2 = api.prox_residual(wk, lk_new, p, q, tk, gam, want=("Axprox",))         # :212
    Fk_new = bk1 * lk_new - ev2["Axprox"] - wlk
    lap("plan", t0)
    return lk_new, Fk_new, {"E": ev["count"], "ms_plan": tm["plan"], "ms_asat": tm["asat"], "ms_amg": tm["amg"], "itamg": itamg, "resamg": resamg, "info": info, "ll": ll, "ls_passes": passes,
                            "nnzH": H0.nnz, "Fk_old_norm": float(torch.linalg.norm(Fk_old)),
                            "Fk_new_norm": float(torch.linalg.norm(Fk_new)), "zeta": zeta}


def ssn_step_host(hstate, amg_options=None):
    """The same step through host buffers: every input is copied host->device (from pinned memory
    when the caller pinned it) and the step's result is read back device->host."""
    import torch
    dev = {k: (v.cuda(non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in hstate.items()}
    lk_new, Fk_new, info = ssn_step(dev, amg_options)
    return lk_new.cpu(), Fk_new.cpu(), info


def capture_state(c, r, l, p, q, gama=np.inf, outer=30, ssn_it=1, warm_maxit=100, run_to_end=False, keep_plans=False):
    """Runs the Class1 solve on the device until SsN step ``ssn_it`` of outer iteration ``outer`` and
    returns the APD state that step reads (a realistic system for benchmarks / parity tests).
    ``run_to_end`` lets the solve finish instead of stopping there and returns ``(state, solve_result)``."""
    import torch
    box = {}

    class _Stop(Exception):
        pass

    def hook(st):
        if not box and st["k"] >= outer and st["ssn_it"] >= ssn_it:
            box.update({"wk": st["wk"].clone(), "lk": st["lk"].clone(), "wlk": st["wlk"].clone(), "bk1": st["bk1"],
                        "tk": st["tk"], "k": st["k"], "ssn_it": st["ssn_it"], "E": st["E"], "ak": st["ak"], "bk": st["bk"]})
            if keep_plans:                                  # sparse at late APD states: tools/save_bench_state.py
                for key in ("xk", "vk"):
                    nz = torch.nonzero(st[key]).reshape(-1)
                    box[key + "_idx"] = nz.cpu().numpy().astype(np.int64); box[key + "_val"] = st[key][nz].cpu().numpy()
            if not run_to_end:
                raise _Stop()
    res = None
    try:
        res = APD_SsN_Class1(c, r, l, p, q, gama, on_ssn_step=hook, warm_maxit=warm_maxit,
                             max_outer=None if run_to_end else outer + 1)
    except _Stop:
        pass
    if not box:
        raise RuntimeError("the solve converged before the requested state")
    box["p"] = _t(p, torch); box["q"] = _t(q, torch); box["gama"] = float(gama) if np.isscalar(gama) else gama
    return (box, res) if run_to_end else box
