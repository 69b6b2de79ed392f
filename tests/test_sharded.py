"""Host logic of the row-sharded SsN step (codes-of-ipd-ssn-amg-method_b200/sharded.py) on CPU:
two gloo ranks, the plan operators supplied by an adapter over the oracle.  The sharded step must
reproduce the unsharded one: same active set / ASAt, same AMG cycle count, same line-search
length, iterates equal to rounding (the column-sum reduction order changes with the world size)."""
import importlib
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

AMG_OPTS = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 0.25, "smoth": 5, "cycle": "w", "isnsp": 1,
            "inter": 1, "guess": None}


class OracleOps:
    """The oracle behind the operator names ShardedStep uses (tests only)."""

    def __init__(self):
        import oracle
        self.o = oracle

    @staticmethod
    def _np(t):
        return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)

    def prox_residual(self, w, lam, p, q, tk, gama, want):
        w, lam, p, q = (self._np(v) for v in (w, lam, p, q))
        z = 1 / tk * (w - self.o.Aty(lam, p, q))
        s = (z >= 0) & (z <= gama)
        px = np.minimum(np.maximum(z, 0.0), gama)
        out = {"norm2": float(px @ px), "count": int(s.sum())}
        if "Axprox" in want:
            out["Axprox"] = torch.from_numpy(self.o.Ax(px, p, q))
        if "s" in want:
            out["s"] = torch.from_numpy(s.astype(np.uint8))
        if "prox" in want:
            out["prox"] = torch.from_numpy(px)
        return out

    def prox_trials(self, w, lamT, p, q, tk, gama):
        return torch.tensor([self.prox_residual(w, l, p, q, tk, gama, ())["norm2"] for l in lamT], dtype=torch.float64)

    def prox_trials_lin(self, w, lam, zeta, p, q, tk, delta, ll0, nt):
        lam, zeta = torch.as_tensor(lam), torch.as_tensor(zeta)
        vals = [self.prox_residual(w, lam + delta ** (ll0 + t) * zeta, p, q, tk, np.inf, ()) for t in range(nt)]
        slots = float(self._np(w).size)                                # the oracle screens nothing: every entry survives
        return torch.tensor([v["norm2"] for v in vals] + [slots], dtype=torch.float64)

    def trial_vectors(self, lam, zeta, wlk, delta, ll0, nt):
        lamT = torch.stack([lam + delta ** (ll0 + t) * zeta for t in range(nt)])
        f0 = torch.stack([v for t in range(nt) for v in (lamT[t] @ lamT[t], wlk @ lamT[t])])
        return lamT, f0

    def active_lin(self, s, m_loc, n, r0, m):
        S = self._np(s).reshape(n, m_loc)                 # column-major slab: S[j, i]
        j, i = np.nonzero(S)                              # slab CSC order
        return torch.from_numpy((i + r0 + j * m).astype(np.int64))

    def asat_from_lin(self, lin_sorted, p, q):
        p, q = self._np(p), self._np(q)
        s = np.zeros(p.size * q.size, dtype=bool)
        s[self._np(lin_sorted)] = True
        return self.o.ASAt(s, p, q)

    def hybrid_amg(self, pd, opts):
        N = pd["p"].numel() + pd["q"].numel()
        d = {"bk1": pd["bk1"], "tk": pd["tk"], "p": self._np(pd["p"]), "q": self._np(pd["q"]),
             "T": sp.diags(np.zeros(N)), "H0": pd["H0"], "z": self._np(pd["z"])}
        zeta, it, res, info = self.o.Hybrid_AMG(d, opts)
        return torch.from_numpy(np.asarray(zeta)), it, res, info

    def hybrid_twogrid(self, pd, opts):
        N = pd["p"].numel() + pd["q"].numel()
        d = {"bk1": pd["bk1"], "tk": pd["tk"], "p": self._np(pd["p"]), "q": self._np(pd["q"]),
             "T": sp.diags(np.zeros(N)), "H0": pd["H0"], "z": self._np(pd["z"])}
        zeta, it, res, info = self.o.Hybrid_twogrid(d, opts)
        return torch.from_numpy(np.asarray(zeta)), it, res, info

    def rng_reset(self):
        self.o.rng_reset()

    # ---- operators of the sharded outer loop (sharded_driver.py)
    def Ax(self, x, p, q):
        return torch.from_numpy(self.o.Ax(self._np(x), self._np(p), self._np(q)))

    def Aty(self, y, p, q):
        return torch.from_numpy(self.o.Aty(self._np(y), self._np(p), self._np(q)))

    def invAAt(self, x, p, q, sg):
        return torch.from_numpy(np.asarray(self.o.invAAt(self._np(x), self._np(p), self._np(q), sg)).reshape(-1))

    def apd_begin(self, c, xk, vk, p, q, ak, bk):
        c, xk, vk = (self._np(v) for v in (c, xk, vk))
        wk = -c + bk * (xk + ak * vk) / ak ** 2                          # Class1/APD_SsN_Class1.m:125
        return torch.from_numpy(wk), self.Ax(xk, p, q)

    def apd_end(self, c, wk, xk, lam, p, q, tk, ak, gama):
        c, wk, xk, lam, p, q = (self._np(v) for v in (c, wk, xk, lam, p, q))
        prox = lambda x: np.minimum(np.maximum(0.0, x), gama)
        aty = self.o.Aty(lam, p, q)
        xk1 = prox((wk - aty) / tk); vk1 = xk1 + (xk1 - xk) / ak         # :239
        d = xk1 - prox(xk1 - c - aty)                                    # :254
        return torch.from_numpy(xk1), torch.from_numpy(vk1), torch.from_numpy(self.o.Ax(xk1, p, q)), float(c @ xk1), float(d @ d)

    def rand(self, count):
        return torch.from_numpy(np.asarray(self.o.rand(count), dtype=np.float64).reshape(-1))

    def warm_stage(self, stage, xk, vk, wk, pik, lk2, dd, c, p, q, b, lk1, axk, y, ak, bk, gk, gama):
        """The two fused stages of a warm-start iteration restated line by line (Class1/warmup_class1.m:59-75),
        updating the torch tensors in place like the CUDA kernels."""
        X, V, W, PI, L2, C_ = (self._np(t) for t in (xk, vk, wk, pik, lk2, c))
        pn, qn, bn = self._np(p), self._np(q), self._np(b)
        muf = 0.0
        bk1 = bk / (1 + ak); etafk = (1 + ak) * gk + muf * ak; sgk = 1 / bk1; etagk = (1 + ak) * bk
        if stage == 0:
            wxk = (ak * gk * V + (gk + muf * ak) * X) / etafk                       # :62
            h1 = self._np(lk1) - (self._np(axk) - bn) / bk                          # :65
            h2 = L2 - (X - W) / bk - (ak / bk) * (PI - W)
            cAw = -self.o.Aty(bn, pn, qn) - W; cAlk = self.o.Aty(h1, pn, qn) + h2   # :66
            d = etafk * wxk - ak ** 2 * (C_ + cAlk + sgk * cAw)                     # :67
            dd.copy_(torch.from_numpy(d))
            return torch.from_numpy(self.o.Ax(d, pn, qn))
        tt = sgk * ak ** 2
        prox = lambda v: np.minimum(np.maximum(0.0, v), gama)
        x1 = (self._np(dd) - self.o.Aty(self._np(y), pn, qn)) / (etafk + tt)        # :70
        v1 = x1 + (x1 - X) / ak
        wwk = (ak * PI + W) / (1 + ak)                                               # :61
        blk2 = L2 + (ak / bk) * (v1 - PI)                                            # :72
        w1 = prox(wwk - (ak ** 2 / etagk) * (-blk2))                                 # :73
        pi1 = w1 + (w1 - W) / ak
        l2 = L2 + (ak / bk) * (v1 - pi1)                                             # :75
        for t, a in ((xk, x1), (vk, v1), (wk, w1), (pik, pi1), (lk2, l2)):
            t.copy_(torch.from_numpy(a))
        return torch.from_numpy(self.o.Ax(v1, pn, qn)), torch.from_numpy(self.o.Ax(x1, pn, qn))


def make_state(m, n, seed):
    rs = np.random.RandomState(seed)
    c = rs.random_sample(m * n)
    w = -c + 0.9 * rs.random_sample(m * n)
    lam = 0.3 * rs.standard_normal(m + n)
    wlk = rs.standard_normal(m + n)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return {"wk": t(w), "lk": t(lam), "wlk": t(wlk), "p": t(np.ones(m)), "q": t(np.ones(n)), "bk1": 0.25, "tk": 0.7,
            "gama": float("inf")}


def _worker(rank, world, port, m, n, seed, outdir):
    import importlib
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded")
    step = sharded.make_sharded_step(make_state(m, n, seed), rank, world, ops=OracleOps(), dist=dist, amg_options=AMG_OPTS)
    lk_new, Fk_new, info = step()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), lk=lk_new.numpy(), Fk=Fk_new.numpy(), E=info["E"], it=info["itamg"],
             ll=info["ll"], nnzH=info["nnzH"], coll=info["collectives"])
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("m,n", [(24, 20), (37, 30)])
def test_sharded_step_matches_unsharded_gloo(m, n):
    import importlib
    import torch.multiprocessing as mp
    sharded = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded")
    seed = 11
    ref = sharded.make_sharded_step(make_state(m, n, seed), 0, 1, ops=OracleOps(), dist=None, amg_options=AMG_OPTS)
    lk_ref, Fk_ref, info_ref = ref()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), m, n, seed, d), nprocs=2, join=True)
        outs = [np.load(os.path.join(d, f"rank{r}.npz")) for r in range(2)]
    for o in outs:
        assert int(o["E"]) == info_ref["E"] and int(o["nnzH"]) == info_ref["nnzH"]
        assert int(o["it"]) == info_ref["itamg"] and int(o["ll"]) == info_ref["ll"]
        assert np.allclose(o["lk"], lk_ref.numpy(), rtol=1e-9, atol=1e-12)
        assert np.allclose(o["Fk"], Fk_ref.numpy(), rtol=1e-8, atol=1e-10)
        assert int(o["coll"]) > 0
    assert np.array_equal(outs[0]["lk"], outs[1]["lk"])          # replicated AMG: identical on every rank


def _grid(g):
    problems = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.problems")
    return problems.grid_problem(g, seed=0)


def _solve_worker(rank, world, port, g, outdir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _run_sharded_solve(rank, world, g, dist)
    np.savez(os.path.join(outdir, f"solve{rank}.npz"), fxk=np.array(res["fxk"]), kx=np.array(res["KKT_xk"]), kl=np.array(res["KKT_lk"]),
             lk=res["lk"].numpy(), xk=res["xk"].numpy(), its=np.array(res["stats"]["ssn_its"]), coll=res["stats"]["collectives"])
    dist.barrier()
    dist.destroy_process_group()


def _run_sharded_solve(rank, world, g, dist, inner_solver=4, fused_warmup=True):
    sd = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded_driver")
    sharded = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded")
    P = _grid(g)
    m, n = P["m"], P["n"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    r0, r1 = sharded.row_range(rank, world, m)
    c_loc = sharded.shard_plan_vector(t(P["c"]), m, n, r0, r1)
    ops = OracleOps(); ops.rng_reset()
    return sd.APD_SsN_Class1_sharded(c_loc, t(P["r"]), t(P["l"]), t(P["p"]), t(P["q"]), rank, world, ops=ops, dist=dist,
                                     amg_options=AMG_OPTS, warm_maxit=30, max_outer=4, inner_solver=inner_solver, fused_warmup=fused_warmup)


def test_sharded_outer_loop_matches_oracle_driver_and_two_ranks():
    """The row-sharded Class 1 solve (warm start + APD outer loop + SsN steps) on one rank against the
    oracle's own driver, and on two gloo ranks against one rank: same number of SsN steps per outer
    iteration, objective / KKT histories and iterates equal to rounding."""
    import torch.multiprocessing as mp
    import oracle
    from oracle import driver as odrv
    g = 5
    one = _run_sharded_solve(0, 1, g, None)
    P = _grid(g)
    oracle.rng_reset()
    calls = {"n": 0}

    class _Stop(Exception):
        pass

    def hook(st):
        calls["n"] += 1
    try:
        ref = odrv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], np.inf, maxit=4, warm_maxit=30, on_ssn_step=hook)
    except TypeError:
        ref = None
    if ref is not None:
        k = len(one["fxk"])
        assert np.allclose(one["fxk"], np.asarray(ref["fxk"])[:k], rtol=1e-9, atol=1e-12)
        assert sum(one["stats"]["ssn_its"]) == calls["n"]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_solve_worker, args=(2, _free_port(), g, d), nprocs=2, join=True)
        outs = [np.load(os.path.join(d, f"solve{r}.npz")) for r in range(2)]
    for o in outs:
        assert np.array_equal(o["its"], np.array(one["stats"]["ssn_its"]))
        assert np.allclose(o["fxk"], one["fxk"], rtol=1e-9, atol=1e-12)
        assert np.allclose(o["kl"], one["KKT_lk"], rtol=1e-6, atol=1e-10) and np.allclose(o["kx"], one["KKT_xk"], rtol=1e-6, atol=1e-10)
        assert np.allclose(o["lk"], one["lk"].numpy(), rtol=1e-8, atol=1e-11)
        assert int(o["coll"]) > 0
    m = g * g
    full = np.concatenate([o["xk"].reshape(m, -1) for o in outs], axis=1)       # slabs are (n, m_loc) row-major
    assert np.allclose(full.reshape(-1), one["xk"].numpy(), rtol=1e-8, atol=1e-11)


def test_fused_and_operator_by_operator_sharded_warm_start_agree():
    """The staged warm start (two fused stages per iteration with the column sums exchanged in between)
    against the operator-by-operator warm start of the same driver."""
    g = 5
    a = _run_sharded_solve(0, 1, g, None, fused_warmup=True)
    b = _run_sharded_solve(0, 1, g, None, fused_warmup=False)
    assert a["stats"]["ssn_its"] == b["stats"]["ssn_its"]
    assert np.allclose(a["fxk"], b["fxk"], rtol=1e-9, atol=1e-11)
    assert np.allclose(a["lk"].numpy(), b["lk"].numpy(), rtol=1e-8, atol=1e-11)


def test_sharded_outer_loop_with_the_two_grid_solver():
    """inner_solver = 5 (Hybrid_twogrid) in the sharded outer loop: the inner systems are solved to the same
    tolerance by a different multilevel method, so the outer iteration follows the inner_solver = 4 run."""
    g = 5
    a = _run_sharded_solve(0, 1, g, None, inner_solver=4)
    b = _run_sharded_solve(0, 1, g, None, inner_solver=5)
    assert a["stats"]["ssn_its"] == b["stats"]["ssn_its"]
    assert np.allclose(a["fxk"], b["fxk"], rtol=1e-7, atol=1e-10)
    assert np.allclose(a["lk"].numpy(), b["lk"].numpy(), rtol=1e-6, atol=1e-9)
    with pytest.raises(ValueError):
        _run_sharded_solve(0, 1, g, None, inner_solver=3)


def test_row_ranges_and_slab_layout():
    import importlib
    sharded = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded")
    m, n, world = 13, 5, 4
    X = np.arange(m * n, dtype=np.float64).reshape(n, m).T           # X[i, j] = i + j*m (column-major vec)
    x = torch.from_numpy(np.ascontiguousarray(X.reshape(-1, order="F")))
    covered = []
    for r in range(world):
        r0, r1 = sharded.row_range(r, world, m)
        covered += list(range(r0, r1))
        slab = sharded.shard_plan_vector(x, m, n, r0, r1).numpy()
        assert np.array_equal(slab.reshape(n, r1 - r0).T, X[r0:r1, :])
    assert covered == list(range(m))
