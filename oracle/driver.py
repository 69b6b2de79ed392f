"""Oracle: the calling scripts, restated so that realistic SsN states and end-to-end results
exist to test against -- reference Class1/APD_SsN_Class1.m, Class1/warmup_class1.m
(test infrastructure only)."""
import time

import numpy as np
import scipy.sparse as sp

from . import rng as _rng
from .pcg import PCG
from .plan_ops import ASAt, Aty, Ax, invAAt
from .solvers import Hybrid_AMG, aug_PCG

CLASS1_AMG_OPTIONS = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 1 / 4, "smoth": 5,
                      "cycle": "w", "isnsp": 1, "inter": 1, "guess": None}   # APD_SsN_Class1.m:87-88
CLASS1_PCG_OPTIONS = {"retol": 1e-11, "maxit": 1e4, "precd": 2, "guess": None}  # :81


def warmup_class1(c, r, l, p, q, gama, res=1e-1, maxit=np.inf):
    """A-ADMM warm start -- reference Class1/warmup_class1.m:18-96 (maxit form only)."""
    if maxit == np.inf:
        maxit = 500
    m, n = l.size, r.size
    prox = lambda x: np.minimum(np.maximum(0.0, x), gama)
    b = np.concatenate([r, l]); Atb = Aty(b, p, q); z0 = np.zeros(m + n)
    muf = 0.0; gk = 1.0; bk = 1.0
    xk = np.zeros(m * n); vk = xk.copy(); wk = xk.copy(); pik = wk.copy()
    lk = np.concatenate([z0, xk])
    for _ in range(int(maxit)):                                         # warmup_class1.m:43-95
        ak = bk; bk1 = bk / (1 + ak)
        gk1 = (gk + muf * ak) / (1 + ak)
        etafk = (1 + ak) * gk + muf * ak
        sgk = 1 / bk1; etagk = (1 + ak) * bk
        wwk = (ak * pik + wk) / (1 + ak)
        wxk = (ak * gk * vk + (gk + muf * ak) * xk) / etafk
        hlk = lk - 1 / bk * np.concatenate([Ax(xk, p, q) - b, xk - wk]) \
            + ak / bk * np.concatenate([z0, -(pik - wk)])
        cAw = -Atb - wk; cAlk = Aty(hlk[:m + n], p, q) + hlk[m + n:]
        dd = etafk * wxk - ak ** 2 * (c + cAlk + sgk * cAw)
        tt = sgk * ak ** 2; sg = 1 + etafk / tt
        xk1 = (dd - Aty(invAAt(Ax(dd, p, q), p, q, sg), p, q)) / (etafk + tt)
        vk1 = xk1 + (xk1 - xk) / ak
        blk = lk + ak / bk * np.concatenate([Ax(vk1, p, q) - b, vk1 - pik])
        wk1 = prox(wwk - ak ** 2 / etagk * (-blk[m + n:]))
        pik1 = wk1 + (wk1 - wk) / ak
        lk1 = lk + ak / bk * np.concatenate([Ax(vk1, p, q) - b, vk1 - pik1])
        gk = gk1; bk = bk1; xk = xk1; vk = vk1; wk = wk1; pik = pik1; lk = lk1
    return xk, lk[:m + n]


def APD_SsN_Class1(c, r, l, p, q, gama, inner_solver=4, maxit=100, KKT_Tol=1e-6,
                   warm_maxit=100, on_ssn_step=None, verbose=False, max_seconds=None, max_outer=None):
    """The APD outer loop with SsN inner loop -- reference Class1/APD_SsN_Class1.m:32-275.

    ``on_ssn_step(state_dict)`` is called before every inner linear solve with everything the
    solve reads (wk, lk_old, wlk, bk1, tk, s, Fk_old), so tests can snapshot realistic systems.
    Returns a dict with the final iterate and histories.
    """
    m, n = l.size, r.size
    prox = lambda x: np.minimum(np.maximum(0.0, x), gama)               # :32
    if np.all(np.isinf(gama)):                                          # prob < 3   :183-184
        ls_term = lambda z: np.linalg.norm(prox(z)) ** 2
    else:                                                               # prob = 3 (capacities)   :185-186
        ls_term = lambda z: np.linalg.norm(z) ** 2 - np.linalg.norm(z - prox(z)) ** 2
    b = np.concatenate([r, l])
    bk = 1.0
    SsN_IT = 50; SsN_Tol1 = 1e-11; nu = 0.2; delta = 0.9; ll_max = 500  # :36
    xk, lk = warmup_class1(c, r, l, p, q, gama, 0, warm_maxit)          # :59
    vk = xk.copy()
    fxk = [float(c @ xk)]
    KKT_lk = [np.linalg.norm(Ax(xk, p, q) - b)]
    KKT_xk = [np.linalg.norm(xk - prox(xk - c - Aty(lk, p, q)))]
    amg_options = dict(CLASS1_AMG_OPTIONS); pcg_options = dict(CLASS1_PCG_OPTIONS)
    stats = {"ssn_its": [], "lin_its": [], "ls_trials": 0, "converged": False, "amg_calls": 0,
             "steps": []}                 # steps: (k, ssn_it, E, ncomp, its, ll, |Fk_new|) per SsN step
    t0 = time.time()
    for k in range(1, maxit + 1):                                       # :101
        resk = max(KKT_xk[k - 1], KKT_lk[k - 1])
        ak = np.sqrt(k ** 2 * bk)                                       # :113
        bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2               # :120
        SsN_Tol = max(bk1 / (k ** 2), SsN_Tol1)                         # :123
        wk = -c + bk * (xk + ak * vk) / ak ** 2                         # :125
        wlk = bk1 * (lk - 1 / bk * (Ax(xk, p, q) - b)) - b              # :126
        ssn_it = 0; lk_new = lk.copy()
        zk = 1 / tk * (wk - Aty(lk_new, p, q))                          # :129
        Fk_new = bk1 * lk_new - Ax(prox(zk), p, q) - wlk                # :130
        Fk_res = np.linalg.norm(Fk_new)
        its = []
        while np.linalg.norm(Fk_new) > SsN_Tol:                         # :137
            ssn_it += 1; lk_old = lk_new
            zk = 1 / tk * (wk - Aty(lk_old, p, q))                      # :139
            s = (zk >= 0) & (zk <= gama); t = np.zeros(m + n)           # :140
            T = sp.diags(t, format="csc"); H0 = ASAt(s, p, q)           # :142
            Fk_old = bk1 * lk_old - Ax(prox(zk), p, q) - wlk            # :144
            if on_ssn_step is not None:
                on_ssn_step({"k": k, "ssn_it": ssn_it, "wk": wk, "lk_old": lk_old, "wlk": wlk,
                             "bk1": bk1, "tk": tk, "s": s, "Fk_old": Fk_old, "H0": H0})
            if inner_solver == 2:                                       # :149-152
                Jk = bk1 * sp.identity(m + n, format="csc") + (T + H0) / tk
                zeta, itpcg, respcg, _ = PCG(Jk, -Fk_old, pcg_options); info = [0, 0]
            else:
                prob_data = {"bk1": bk1, "tk": tk, "q": q, "p": p, "T": T, "H0": H0, "z": -Fk_old}
                if inner_solver == 3:
                    zeta, itpcg, respcg, info = aug_PCG(prob_data, pcg_options)
                else:
                    zeta, itpcg, respcg, info = Hybrid_AMG(prob_data, amg_options)
                    stats["amg_calls"] += 1
            its.append(itpcg)
            f0 = bk1 / 2 * np.linalg.norm(lk_old) ** 2 - wlk @ lk_old   # :182
            cFk_old = f0 + 0.5 * tk * ls_term(zk)
            ll = 0; lk_new = lk_old + delta ** ll * zeta
            f0 = bk1 / 2 * np.linalg.norm(lk_new) ** 2 - wlk @ lk_new
            zk = 1 / tk * (wk - Aty(lk_new, p, q))
            cFk_new = f0 + 0.5 * tk * ls_term(zk)
            ress = abs(Fk_old @ zeta)
            while cFk_new > cFk_old - nu * delta ** ll * ress:          # :199-211
                ll += 1; lk_new = lk_old + delta ** ll * zeta
                f0 = bk1 / 2 * np.linalg.norm(lk_new) ** 2 - wlk @ lk_new
                zk = 1 / tk * (wk - Aty(lk_new, p, q))
                cFk_new = f0 + 0.5 * tk * ls_term(zk)
                if ll == ll_max:
                    break
            stats["ls_trials"] += ll + 1
            Fk_new = bk1 * lk_new - Ax(prox(zk), p, q) - wlk            # :212
            nFn = np.linalg.norm(Fk_new)
            stats["steps"].append((k, ssn_it, int(np.count_nonzero(s)), int(info[0]), int(itpcg), ll, float(nFn)))
            if verbose:
                print(f"   SsN: it={ssn_it:3d} |Fk|={nFn:.2e} ll={ll:3d} info={list(info)} "
                      f"its={itpcg} res={respcg:.2e}")
            if nFn <= SsN_Tol:
                break
            if abs(np.linalg.norm(Fk_old) - nFn) < SsN_Tol / 100:       # :219
                break
            if ssn_it == SsN_IT:
                break
            if Fk_res / nFn >= 2:
                Fk_res = nFn
        lk1 = lk_new; xk1 = prox(zk); vk1 = xk1 + (xk1 - xk) / ak       # :239
        kl = np.linalg.norm(Ax(xk1, p, q) - b)
        kx = np.linalg.norm(xk1 - prox(xk1 - c - Aty(lk1, p, q)))
        rr = [kx / (1 + KKT_xk[0]), kl / (1 + KKT_lk[0])]
        if bk1 < 1e-8 and max(rr) > resk:                               # :245-249
            xk1 = xk; lk1 = lk; vk1 = xk; bk1 = float(_rng.rand(1)[0])
        bk = bk1; xk = xk1; lk = lk1; vk = vk1                          # :251
        fxk.append(float(c @ xk)); KKT_lk.append(np.linalg.norm(Ax(xk, p, q) - b))
        KKT_xk.append(np.linalg.norm(xk - prox(xk - c - Aty(lk, p, q))))
        stats["ssn_its"].append(ssn_it); stats["lin_its"].append(its)
        rr = [KKT_xk[k] / (1 + KKT_xk[0]), KKT_lk[k] / (1 + KKT_lk[0])]
        if verbose:
            print(f"APD: it={k:3d} KKT(xk)={rr[0]:.2e} KKT(lk)={rr[1]:.2e} fk={fxk[-1]:.8e} "
                  f"t={time.time() - t0:.1f}s")
        if max(rr) <= KKT_Tol:                                          # :266
            stats["converged"] = True
            break
        if max_outer is not None and k >= max_outer:
            break
        if max_seconds is not None and time.time() - t0 > max_seconds:
            break
    return {"xk": xk, "lk": lk, "fxk": fxk, "KKT_xk": KKT_xk, "KKT_lk": KKT_lk, "outer_its": k,
            "rel_kkt": max(rr), "stats": stats, "seconds": time.time() - t0}


# ------------------------------------------------------------------ Class 2: partial optimal transport

CLASS2_AMG_OPTIONS = {"retol": 1e-11, "bigph": 1, "maxit": 40, "theta": 1 / 4, "smoth": 10,
                      "cycle": "w", "isnsp": 1, "inter": 1, "guess": None}   # APD_SsN_Class2.m:80-81


def warmup_class2(c, r, l, p, q, mu, phi, res=1e-1, maxit=np.inf):
    """A-ADMM warm start for partial OT -- reference Class2/warmup_class2.m:2-108 (maxit form)."""
    from .plan_ops import invHHt
    if maxit == np.inf:
        maxit = 500
    m, n = l.size, r.size; N = m + n; mn = m * n
    prox = lambda x: np.maximum(0.0, x)                                  # :21
    b = np.concatenate([r, l, [mu]])
    Htb = np.concatenate([Aty(b[:N], p, q) + b[-1] * phi, b[:N]])        # :22
    wc = np.concatenate([c, np.zeros(N)]); z0 = np.zeros(N + 1)          # :23
    muf = 0.0; gk = 1.0; bk = 1.0
    uk = np.zeros(mn + N); vk = uk.copy(); wk = uk.copy(); pik = wk.copy()
    lk = np.concatenate([z0, uk])                                       # :26
    xk, yk, zk = uk[:mn], uk[mn:mn + n], uk[mn + n:]
    for _ in range(int(maxit)):                                         # :46-107
        ak = bk; bk1 = bk / (1 + ak)
        gk1 = (gk + muf * ak) / (1 + ak)
        etafk = (1 + ak) * gk + muf * ak
        sgk = 1 / bk1; etagk = (1 + ak) * bk
        wwk = (ak * pik + wk) / (1 + ak)
        wuk = (ak * gk * vk + (gk + muf * ak) * uk) / etafk
        Hu = np.concatenate([Ax(xk, p, q) + np.concatenate([yk, zk]), [phi @ xk]])
        hlk = lk - 1 / bk * np.concatenate([Hu - b, uk - wk]) + ak / bk * np.concatenate([z0, -(pik - wk)])   # :66
        cAw = -Htb - wk; cAlk = hlk[N + 1:]
        cAlk = cAlk + np.concatenate([Aty(hlk[:N], p, q) + hlk[N] * phi, hlk[:N]])    # :68
        dd = etafk * wuk - ak ** 2 * (wc + cAlk + sgk * cAw)             # :69
        tt = sgk * ak ** 2; sg = 1 + etafk / tt
        Hdd = np.concatenate([Ax(dd[:mn], p, q) + dd[mn:], [phi @ dd[:mn]]])          # :71
        ff = invHHt(Hdd, p, q, sg, phi)                                  # :72
        uk1 = (dd - np.concatenate([Aty(ff[:N], p, q) + ff[-1] * phi, ff[:N]])) / (etafk + tt)   # :73
        vk1 = uk1 + (uk1 - uk) / ak
        b0 = np.concatenate([Ax(vk1[:mn], p, q) + vk1[mn:], [phi @ vk1[:mn]]]) - b     # :75
        blk = lk + ak / bk * np.concatenate([b0, vk1 - pik])
        wk1 = prox(wwk - ak ** 2 / etagk * (-blk[N + 1:]))               # :77
        pik1 = wk1 + (wk1 - wk) / ak
        lk1 = lk + ak / bk * np.concatenate([b0, vk1 - pik1])            # :79
        gk = gk1; bk = bk1; uk = uk1; vk = vk1; wk = wk1; pik = pik1; lk = lk1
        xk, yk, zk = uk[:mn], uk[mn:mn + n], uk[mn + n:]
    return uk, lk[:N + 1]


def APD_SsN_Class2(c, r, l, p, q, mu, phi, inner_solver=4, maxit=100, KKT_Tol=1e-6, warm_maxit=100,
                   on_ssn_step=None, verbose=False, max_seconds=None, max_outer=None):
    """APD + SsN for partial OT -- reference Class2/APD_SsN_Class2.m:25-285 (inner solvers 3 =
    PCG4POT, 4 = AMG4POT).  Unknowns u = [x (mn); y (n); z (m)], duals lk (n+m+1)."""
    from .solvers import AMG4POT, PCG4POT
    m, n = l.size, r.size; N = m + n; mn = m * n
    prox = lambda x: np.maximum(0.0, x)                                  # :25
    b = np.concatenate([r, l, [mu]]); wc = np.concatenate([c, np.zeros(N)])
    bk = 1.0
    SsN_IT = 50; SsN_Tol1 = 1e-10; nu = 0.2; delta = 0.9; ll_max = 500  # :28
    uk, lk = warmup_class2(c, r, l, p, q, mu, phi, 0, warm_maxit)       # :50
    vk = uk.copy()
    Hmul = lambda u: np.concatenate([Ax(u[:mn], p, q) + u[mn:], [phi @ u[:mn]]])       # H*u, H = [G IY IZ]
    Htmul = lambda lam: np.concatenate([Aty(lam[:N], p, q) + lam[N] * phi, lam[:N]])    # H'*lam

    def kkts(u, lam):
        x, y, z = u[:mn], u[mn:mn + n], u[mn + n:]
        return (np.linalg.norm(x - np.maximum(x - c - (Aty(lam[:N], p, q) + lam[N] * phi), 0)),
                np.linalg.norm(y - np.maximum(y - lam[:n], 0)),
                np.linalg.norm(z - np.maximum(z - lam[n:N], 0)),
                np.linalg.norm(Hmul(u) - b))

    k0 = kkts(uk, lk)
    fxk = [float(c @ uk[:mn])]; KKT = [k0]
    amg_options = dict(CLASS2_AMG_OPTIONS); pcg_options = dict(CLASS1_PCG_OPTIONS)
    stats = {"ssn_its": [], "lin_its": [], "ls_trials": 0, "converged": False, "amg_calls": 0, "steps": []}
    t0 = time.time()
    rr = [np.inf]
    for k in range(1, maxit + 1):                                       # :95
        resk = max(KKT[k - 1])
        ak = np.sqrt(k ** 2 * bk)                                       # :116
        bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2
        SsN_Tol = max(bk1 / (k ** 2), SsN_Tol1)                         # :119
        wk = -wc + bk * (uk + ak * vk) / ak ** 2                        # :121
        wlk = bk1 * (lk - 1 / bk * (Hmul(uk) - b)) - b                  # :122
        ssn_it = 0; lk_new = lk.copy()
        zk = 1 / tk * (wk - Htmul(lk_new)); pzk = prox(zk)              # :127
        Fk_new = bk1 * lk_new - Hmul(pzk) - wlk                         # :130
        Fk_res = np.linalg.norm(Fk_new)
        its = []
        while np.linalg.norm(Fk_new) > SsN_Tol:                         # :136
            ssn_it += 1; lk_old = lk_new
            zk = 1 / tk * (wk - Htmul(lk_old))                          # :139
            s = zk[:mn] >= 0; t = (zk[mn:] >= 0).astype(np.float64)
            T = sp.diags(t, format="csc"); H0 = ASAt(s, p, q)           # :146
            pzk = prox(zk)
            Fk_old = bk1 * lk_old - Hmul(pzk) - wlk                     # :150
            prob_data = {"bk1": bk1, "tk": tk, "q": q, "p": p, "s": s, "T": T, "H0": H0, "z": -Fk_old, "phi": phi}
            if on_ssn_step is not None:
                on_ssn_step(dict(prob_data, k=k, ssn_it=ssn_it))
            if inner_solver == 3:
                zeta, itpcg, respcg, info = PCG4POT(prob_data, pcg_options)            # :168
            else:
                zeta, itpcg, respcg, info = AMG4POT(prob_data, amg_options, "amg" if inner_solver == 4 else "twogrid")   # :171 / :182
                stats["amg_calls"] += 1
            its.append(itpcg)
            f0 = bk1 / 2 * np.linalg.norm(lk_old) ** 2 - wlk @ lk_old   # :196
            cFk_old = f0 + 0.5 * tk * np.linalg.norm(prox(zk)) ** 2
            ll = 0; lk_new = lk_old + delta ** ll * zeta
            f0 = bk1 / 2 * np.linalg.norm(lk_new) ** 2 - wlk @ lk_new
            zk = 1 / tk * (wk - Htmul(lk_new)); cFk_new = f0 + 0.5 * tk * np.linalg.norm(prox(zk)) ** 2
            ress = abs(Fk_old @ zeta)
            while cFk_new > cFk_old - nu * delta ** ll * ress:          # :205-213
                ll += 1; lk_new = lk_old + delta ** ll * zeta
                f0 = bk1 / 2 * np.linalg.norm(lk_new) ** 2 - wlk @ lk_new
                zk = 1 / tk * (wk - Htmul(lk_new)); cFk_new = f0 + 0.5 * tk * np.linalg.norm(prox(zk)) ** 2
                if ll == ll_max:
                    break
            stats["ls_trials"] += ll + 1
            pzk = prox(zk)
            Fk_new = bk1 * lk_new - Hmul(pzk) - wlk                     # :217
            nFn = np.linalg.norm(Fk_new)
            stats["steps"].append((k, ssn_it, int(np.count_nonzero(s)), int(info[0]), int(itpcg), ll, float(nFn)))
            if verbose:
                print(f"   SsN: it={ssn_it:3d} |Fk|={nFn:.2e} ll={ll:3d} info={list(info)} its={itpcg} res={respcg:.2e}")
            if nFn <= SsN_Tol:
                break
            if abs(np.linalg.norm(Fk_old) - nFn) < SsN_Tol:             # :224 (no /100 in Class2)
                break
            if ssn_it == SsN_IT:
                break
            if Fk_res / nFn >= 2:
                Fk_res = nFn
        lk1 = lk_new; uk1 = prox(zk); vk1 = uk1 + (uk1 - uk) / ak       # :244
        kk = kkts(uk1, lk1)
        rr = [kk[i] / (1 + KKT[0][i]) for i in range(4)]
        if bk1 < 1e-8 and max(rr) > resk:                               # :253-257
            uk1 = uk; lk1 = lk; vk1 = uk; bk1 = 10 * bk1
        bk = bk1; uk = uk1; lk = lk1; vk = vk1                          # :259
        fxk.append(float(c @ uk[:mn])); KKT.append(kkts(uk, lk))
        stats["ssn_its"].append(ssn_it); stats["lin_its"].append(its)
        rr = [KKT[k][i] / (1 + KKT[0][i]) for i in range(4)]
        if verbose:
            print(f"APD: it={k:3d} KKT(x,y,z,l)={['%.2e' % v for v in rr]} fk={fxk[-1]:.8e} t={time.time() - t0:.1f}s")
        if max(rr) <= KKT_Tol:                                          # :274
            stats["converged"] = True
            break
        if max_outer is not None and k >= max_outer:
            break
        if max_seconds is not None and time.time() - t0 > max_seconds:
            break
    return {"uk": uk, "xk": uk[:mn], "lk": lk, "fxk": fxk, "KKT": KKT, "outer_its": k, "rel_kkt": max(rr),
            "stats": stats, "seconds": time.time() - t0}

