"""Oracle: ONE semismooth-Newton step of Class1/APD_SsN_Class1.m:137-212 at a given APD state, at FULL size, on the
host cores -- the CPU arm of bench.py (``--impl reference`` and the ``cpu_baseline`` leg; test infrastructure, never
the product path).

Nothing is sampled or extrapolated: the whole m x n plan, every line-search trial the step takes.  The plan-wide
expressions are the oracle's own ``Aty`` / ``Ax`` / prox / norm restatements (``oracle/plan_ops.py``), evaluated on
contiguous COLUMN BLOCKS of the plan, one block per host thread at the same time (NumPy releases the GIL inside its
kernels).  The operators are separable by columns -- ``Aty`` on columns [j0, j1) needs ``y1[j0:j1]``, ``y2`` and
``q[j0:j1]``; ``Ax`` returns that block's column sums and a partial of the row sums -- so this is the arithmetic of
the reference, expression by expression with the same dense temporaries per block, at the throughput a multi-threaded
element-wise runtime such as MATLAB's gets out of the socket.  ``ASAt`` and ``Hybrid_AMG`` are the oracle's SciPy
restatements on the full system (single-threaded, like MATLAB's sparse kernels).

State fixture (``tests/golden/bench_state_g128_k30.npz``, written on a B200 by ``tools/save_bench_state.py``): the
plans ``xk`` and ``vk`` of the APD iteration are sparse at the benchmarked state, so the fixture holds their nonzeros,
the duals and the scalars; ``wk = -c + bk*(xk + ak*vk)/ak^2`` (:125) and ``wlk`` (:126) are rebuilt here from the
generator's cost and marginals.
"""
import concurrent.futures as cf
import os
import time

import numpy as np
import scipy.sparse as sp

from . import rng as _rng
from .plan_ops import ASAt, Aty, Ax
from .solvers import Hybrid_AMG
from .driver import CLASS1_AMG_OPTIONS


def state_from_fixture(path, problem):
    """-> dict(wk, lk, wlk, p, q, tk, bk1, m, n, k, ssn_it, expect) from the fixture and the generator's
    ``problem`` (``c, r, l, p, q``).  Class1/APD_SsN_Class1.m:113-126."""
    d = np.load(path)
    m, n = int(d["m"]), int(d["n"])
    ak, bk = float(d["ak"]), float(d["bk"])
    bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2                   # :120
    assert bk1 == float(d["bk1"]) and tk == float(d["tk"]), "fixture scalars do not follow :120"
    c = problem["c"]; b = np.concatenate([problem["r"], problem["l"]])
    xk = np.zeros(m * n); xk[d["xk_idx"]] = d["xk_val"]
    wk = xk.copy(); wk[d["vk_idx"]] += ak * d["vk_val"]                 # xk + ak*vk
    wk *= bk / ak ** 2; wk -= c                                          # :125  wk = -c + bk*(xk+ak*vk)/ak^2
    lk = np.array(d["lk"])
    wlk = bk1 * (lk - 1 / bk * (Ax(xk, problem["p"], problem["q"]) - b)) - b   # :126
    expect = {k[7:]: d[k] for k in d.files if k.startswith("expect_")}
    return {"wk": wk, "lk": lk, "wlk": wlk, "p": problem["p"], "q": problem["q"], "tk": tk, "bk1": bk1, "m": m, "n": n,
            "k": int(d["k"]), "ssn_it": int(d["ssn_it"]), "expect": expect}


def ssn_step(state, threads=None, amg_options=None, ll_max=500):
    """One SsN step (:137-212) on the host.  Returns ``(lk_new, Fk_new, info)``; ``info`` has the step's E, nnz(H0),
    component count, W-cycle count, accepted ll and the seconds per phase."""
    wk, lk, wlk, p, q = state["wk"], state["lk"], state["wlk"], state["p"], state["q"]
    tk, bk1, m, n = state["tk"], state["bk1"], state["m"], state["n"]
    nu, delta = 0.2, 0.9                                                # :36
    threads = int(threads or os.cpu_count() or 1)
    nb = min(threads, n)
    edges = np.linspace(0, n, nb + 1).astype(np.int64)
    blocks = [(int(edges[i]), int(edges[i + 1])) for i in range(nb) if edges[i + 1] > edges[i]]
    pool = cf.ThreadPoolExecutor(max_workers=threads)
    tm = {}

    def z_block(lam, j0, j1):                                           # zk = 1/tk*(wk - Aty(lk,p,q)) on columns [j0,j1)   :139
        y = np.concatenate([lam[j0:j1], lam[n:]])
        return 1 / tk * (wk[j0 * m:j1 * m] - Aty(y, p, q[j0:j1]))

    def residual(lam, want_s):                                          # :139-144 / :212
        def work(blk):
            j0, j1 = blk
            z = z_block(lam, j0, j1)
            s = ((z >= 0) & (z <= np.inf)) if want_s else None          # :140
            px = np.maximum(0.0, z)                                      # prox, gama = Inf   :32
            a = Ax(px, p, q[j0:j1])                                      # [column sums of the block ; partial row sums]
            return s, a[:j1 - j0], a[j1 - j0:], float(px @ px)
        out = list(pool.map(work, blocks))
        cols = np.concatenate([o[1] for o in out]); rows = np.sum([o[2] for o in out], axis=0)
        s = np.concatenate([o[0] for o in out]) if want_s else None
        return s, np.concatenate([cols, rows]), float(sum(o[3] for o in out))

    def trial_norm2(lam):                                               # norm(prox(zk))^2 of one Armijo trial   :191-193
        def work(blk):
            px = np.maximum(0.0, z_block(lam, *blk))
            return float(px @ px)
        return float(sum(pool.map(work, blocks)))

    t0 = time.perf_counter()
    s, Axp, n2_old = residual(lk, True)
    Fk_old = bk1 * lk - Axp - wlk                                       # :144
    tm["residual_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    H0 = ASAt(s, p, q)                                                  # :142
    E = int(np.count_nonzero(s)); del s
    tm["asat_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    pd = {"bk1": bk1, "tk": tk, "q": q, "p": p, "T": sp.diags(np.zeros(m + n), format="csc"), "H0": H0, "z": -Fk_old}
    zeta, itamg, resamg, info = Hybrid_AMG(pd, amg_options or CLASS1_AMG_OPTIONS)   # :161
    tm["hybrid_amg_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    f0 = bk1 / 2 * np.linalg.norm(lk) ** 2 - wlk @ lk                   # :182
    cFk_old = f0 + 0.5 * tk * n2_old
    ress = abs(Fk_old @ zeta)
    ll = 0
    while True:                                                         # :189-211
        lk_new = lk + delta ** ll * zeta
        f0 = bk1 / 2 * np.linalg.norm(lk_new) ** 2 - wlk @ lk_new
        cFk_new = f0 + 0.5 * tk * trial_norm2(lk_new)
        if not (cFk_new > cFk_old - nu * delta ** ll * ress) or ll == ll_max:
            break
        ll += 1
    tm["line_search_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    _, Axp, _ = residual(lk_new, False)
    Fk_new = bk1 * lk_new - Axp - wlk                                   # :212
    tm["new_residual_s"] = time.perf_counter() - t0
    pool.shutdown()
    return lk_new, Fk_new, {"E": E, "nnzH": int(H0.nnz), "components": int(info[0]), "itamg": int(itamg), "resamg": float(resamg),
                            "ll": int(ll), "threads": threads, "phases_s": tm, "Fk_old_norm": float(np.linalg.norm(Fk_old)),
                            "Fk_new_norm": float(np.linalg.norm(Fk_new)), "zeta": zeta}


def timed_step(state, threads=None):
    """rng reset (the stream position of the benchmarked step does not matter for its cost) + one timed step."""
    _rng.rng_reset()
    t0 = time.perf_counter()
    lk_new, Fk_new, info = ssn_step(state, threads)
    return (time.perf_counter() - t0) * 1e3, lk_new, Fk_new, info


# ------------------------------------------------------------------ Class 2 (partial OT), config 3 of BASELINE.json

def class2_trivial_state(problem):
    """The APD state of outer iteration 1 from the trivial start ``uk = vk = 0, lk = 0, bk = 1`` (``warm_maxit = 0``, the
    common start of tests/golden/trace_class2_grid64_nowarm_outer3.npz): Class2/APD_SsN_Class2.m:116-122."""
    c, r, l = problem["c"], problem["r"], problem["l"]
    m, n = l.size, r.size
    b = np.concatenate([r, l, [problem["mu"]]])
    ak = 1.0; bk = 1.0
    bk1 = bk / (1 + ak); tk = bk * (1 + ak) / ak ** 2                   # :117
    wk = -np.concatenate([c, np.zeros(m + n)])                          # :121 with uk = vk = 0
    wlk = bk1 * (np.zeros(m + n + 1) - 1 / bk * (0.0 - b)) - b          # :122
    return {"wk": wk, "lk": np.zeros(m + n + 1), "wlk": wlk, "p": problem["p"], "q": problem["q"], "phi": problem["phi"],
            "tk": tk, "bk1": bk1, "m": m, "n": n, "k": 1}


def ssn_step_class2(state, amg_options=None, ll_max=500):
    """One SsN step of Class2/APD_SsN_Class2.m:137-217 on the host (whole-vector NumPy expressions, like the reference's
    MATLAB lines; AMG4POT = the oracle's SciPy restatement).  Returns ``(lk_new, Fk_new, info)``."""
    from .solvers import AMG4POT
    from .driver import CLASS2_AMG_OPTIONS
    wk, lk_old, wlk, p, q, phi = state["wk"], state["lk"], state["wlk"], state["p"], state["q"], state["phi"]
    tk, bk1, m, n = state["tk"], state["bk1"], state["m"], state["n"]
    N = m + n; mn = m * n
    nu, delta = 0.2, 0.9
    prox = lambda x: np.maximum(0.0, x)
    Hmul = lambda u: np.concatenate([Ax(u[:mn], p, q) + u[mn:], [phi @ u[:mn]]])
    Htmul = lambda lam: np.concatenate([Aty(lam[:N], p, q) + lam[N] * phi, lam[:N]])
    tm = {}
    t0 = time.perf_counter()
    zk = 1 / tk * (wk - Htmul(lk_old))                                  # :139
    s = zk[:mn] >= 0; t = (zk[mn:] >= 0).astype(np.float64)
    pzk = prox(zk)
    Fk_old = bk1 * lk_old - Hmul(pzk) - wlk                             # :150
    tm["residual_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    T = sp.diags(t, format="csc"); H0 = ASAt(s, p, q)                   # :146
    tm["asat_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    pd = {"bk1": bk1, "tk": tk, "q": q, "p": p, "s": s, "T": T, "H0": H0, "z": -Fk_old, "phi": phi}
    zeta, itamg, resamg, info = AMG4POT(pd, amg_options or CLASS2_AMG_OPTIONS, "amg")     # :171
    tm["amg4pot_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    f0 = bk1 / 2 * np.linalg.norm(lk_old) ** 2 - wlk @ lk_old           # :196
    cFk_old = f0 + 0.5 * tk * np.linalg.norm(pzk) ** 2
    ress = abs(Fk_old @ zeta)
    ll = 0
    while True:                                                         # :199-213
        lk_new = lk_old + delta ** ll * zeta
        f0 = bk1 / 2 * np.linalg.norm(lk_new) ** 2 - wlk @ lk_new
        zk = 1 / tk * (wk - Htmul(lk_new)); pzk = prox(zk)
        if not (f0 + 0.5 * tk * np.linalg.norm(pzk) ** 2 > cFk_old - nu * delta ** ll * ress) or ll == ll_max:
            break
        ll += 1
    tm["line_search_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    Fk_new = bk1 * lk_new - Hmul(pzk) - wlk                             # :217
    tm["new_residual_s"] = time.perf_counter() - t0
    return lk_new, Fk_new, {"E": int(np.count_nonzero(s)), "nnzH": int(H0.nnz), "components": int(info[0]), "itamg": int(itamg),
                            "resamg": float(resamg), "ll": int(ll), "threads": 1, "phases_s": tm,
                            "Fk_old_norm": float(np.linalg.norm(Fk_old)), "Fk_new_norm": float(np.linalg.norm(Fk_new))}


def timed_step_class2(problem, warm_steps=1):
    """State of SsN step ``warm_steps + 1`` of outer iteration 1 from the trivial start (the first ``warm_steps`` steps are
    run untimed to get there), then ONE timed step.  Returns ``(ms, lk_new, Fk_new, info, seconds of the state build)``."""
    _rng.rng_reset()
    st = class2_trivial_state(problem)
    t0 = time.perf_counter()
    for _ in range(warm_steps):
        st["lk"], _, _ = ssn_step_class2(st)
    t_state = time.perf_counter() - t0
    t0 = time.perf_counter()
    lk_new, Fk_new, info = ssn_step_class2(st)
    return (time.perf_counter() - t0) * 1e3, lk_new, Fk_new, info, t_state
