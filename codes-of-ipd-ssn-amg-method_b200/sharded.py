"""Row-sharded SsN step: one process per GPU, the m x n plan split by rows.

Rank g owns rows ``[g*m/G, (g+1)*m/G)`` of every plan-sized vector as a local column-major
``m_loc x n`` slab.  ``Aty``, prox and the active set are slab-local; the q-weighted row sums are
owner-local and all-gathered (m doubles in total); the p-weighted column sums need the path's one
real exchange, an NCCL ``all_reduce`` of n doubles in which the squared norm and the active count
ride along; the compacted active set crosses NVLink as O(E) global linear indices; the
(m+n)-sized AMG hierarchy is replicated (deterministic, so every rank computes the same zeta and
no broadcast is needed).

The collectives go through ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests); the
plan operators come from ``ops`` -- the CUDA operators of ``api.py`` by default.  The tests pass an
adapter over the CPU oracle instead, which exercises exactly this host logic without a GPU.
"""
import numpy as np


def row_range(rank, world, m):
    return (rank * m) // world, ((rank + 1) * m) // world


def shard_plan_vector(x, m, n, r0, r1):
    """Rows [r0, r1) of a column-major m x n plan vector as a column-major slab."""
    return x.reshape(n, m)[:, r0:r1].contiguous().reshape(-1)


class _CudaOps:
    """The product operators (api.py) under the names the sharded step uses."""

    def __init__(self):
        from . import api
        self.api = api

    def prox_residual(self, w, lam, p, q, tk, gama, want):
        return self.api.prox_residual(w, lam, p, q, tk, gama, want=want)

    def prox_residual_dev(self, w, lam, p, q, tk, gama, want, scal_dev):
        """the same with the norm term and the count left on the device (``scal_dev``): no host read"""
        return self.api.prox_residual(w, lam, p, q, tk, gama, want=want, scal_dev=scal_dev)

    def prox_trials(self, w, lamT, p, q, tk, gama):
        return self.api.prox_trials(w, lamT, p, q, tk, gama)

    def prox_trials_lin(self, w, lam, zeta, p, q, tk, delta, ll0, nt):
        return self.api.prox_trials_lin(w, lam, zeta, p, q, tk, delta, ll0, nt)

    def trial_vectors(self, lam, zeta, wlk, delta, ll0, nt):
        return self.api.trial_vectors(lam, zeta, wlk, delta, ll0, nt)

    def active_lin(self, s, m_loc, n, r0, m):
        return self.api.active_coo(s, m_loc, n, r0, m)

    def asat_from_lin(self, lin_sorted, p, q):
        return self.api.ASAt_coo(lin_sorted, p, q)

    def hybrid_amg(self, prob_data, opts):
        return self.api.Hybrid_AMG(prob_data, opts)

    def hybrid_twogrid(self, prob_data, opts):
        return self.api.Hybrid_twogrid(prob_data, opts)

    def rng_reset(self):
        self.api.rng_reset()

    # ---- the operators the sharded outer loop (sharded_driver.py) adds
    def Ax(self, x, p, q):
        return self.api.Ax(x, p, q)

    def Aty(self, y, p, q):
        return self.api.Aty(y, p, q)

    def invAAt(self, x, p, q, sg):
        return self.api.invAAt(x, p, q, sg)

    def apd_begin(self, c, xk, vk, p, q, ak, bk):
        return self.api.apd_begin(c, xk, vk, p, q, ak, bk)

    def apd_end(self, c, wk, xk, lam, p, q, tk, ak, gama):
        return self.api.apd_end(c, wk, xk, lam, p, q, tk, ak, gama)

    def rand(self, count):
        return self.api.rand(count)

    def warm_stage(self, stage, xk, vk, wk, pik, lk2, dd, c, p, q, b, lk1, axk, y, ak, bk, gk, gama):
        return self.api.warm_stage(stage, xk, vk, wk, pik, lk2, dd, c, p, q, b, lk1, axk, y, ak, bk, gk, gama)


class ShardedStep:
    """One semismooth-Newton step (Class1/APD_SsN_Class1.m:137-212) on a row-sharded plan."""

    def __init__(self, state, rank, world, ops=None, dist=None, amg_options=None, already_sharded=False, inner_solver=4):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.ops = ops if ops is not None else _CudaOps()
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        p, q = state["p"], state["q"]
        self.m, self.n = m, n = p.numel(), q.numel()
        self.r0, self.r1 = r0, r1 = row_range(rank, world, m)
        self.m_loc = r1 - r0
        self.p, self.q = p, q
        self.p_loc = p[r0:r1].contiguous()
        self.w_loc = state["wk"] if already_sharded else shard_plan_vector(state["wk"], m, n, r0, r1)
        self.lk, self.wlk = state["lk"], state["wlk"]
        self.bk1, self.tk, self.gama = state["bk1"], state["tk"], state.get("gama", float("inf"))
        if amg_options is None:
            from .driver import CLASS1_AMG_OPTIONS as amg_options
        self.amg_options = amg_options
        self.counts = [row_range(g, world, m)[1] - row_range(g, world, m)[0] for g in range(world)]
        self.collectives = 0
        self.screen = True
        self.gather_cap_limit = 1 << 20                            # above this many active entries: sizes first, then payloads
        self.profile = False                                       # True: device-synchronised phase times (ms_plan / ms_asat / ms_amg)
        self.last_density = None                                   # share of entries that survived the last screened batch
        self.last_ll = None                                        # backtracking steps the last line search accepted at
        if inner_solver not in (4, 5):
            raise ValueError("inner_solver must be 4 (Hybrid_AMG) or 5 (Hybrid_twogrid)")
        self.inner_solver = inner_solver

    # ---- slab-local view of a dual vector [column part (n) ; row part (m)]
    def _lam_loc(self, lam):
        return self.torch.cat([lam[: self.n], lam[self.n + self.r0: self.n + self.r1]])

    def _all_reduce(self, t):
        if self.world > 1:
            self.dist.all_reduce(t)
            self.collectives += 1
        return t

    def _all_gather_rows(self, rows_loc):
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

    def _gather_active(self, v, E):
        """The ascending union of the ranks' active linear indices (``E`` of them in all, known from the residual's
        all_reduce) with ONE collective and no host round trip: every rank contributes a fixed-capacity block padded
        with INT64_MAX -- capacity = E rounded up to a power of two, which no rank's share can exceed -- and the sorted
        gather's first ``E`` entries are ``find(s)``.  Early SsN steps with millions of active entries (the block would
        be ``world`` times larger than the payload) take the two-phase path."""
        torch = self.torch
        if self.world == 1:
            return v
        if E > self.gather_cap_limit or v.numel() > E:
            return torch.sort(self._all_gather_var(v))[0]
        cap = 1 << max(10, int(E - 1).bit_length())
        pad = torch.full((cap,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=v.device)
        pad[: v.numel()] = v
        out = torch.empty(self.world * cap, dtype=torch.int64, device=v.device)
        if hasattr(self.dist, "all_gather_into_tensor"):
            self.dist.all_gather_into_tensor(out, pad)
        else:
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            self.dist.all_gather(parts, pad); out = torch.cat(parts)
        self.collectives += 1
        return torch.sort(out)[0][:E]

    def _all_gather_var(self, v):
        """all-gather of int64 vectors of different lengths (sizes first, then padded payloads)."""
        if self.world == 1:
            return v
        torch = self.torch
        sz = torch.tensor([v.numel()], dtype=torch.int64, device=v.device)
        sizes = [torch.empty_like(sz) for _ in range(self.world)]
        self.dist.all_gather(sizes, sz)
        sizes = [int(s.item()) for s in sizes]
        mx = max(max(sizes), 1)
        pad = torch.zeros(mx, dtype=v.dtype, device=v.device)
        pad[: v.numel()] = v
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad)
        self.collectives += 2
        return torch.cat([o[:c] for o, c in zip(out, sizes)])

    def residual(self, lam, want_s):
        """Ax(prox(z)) (global, n+m), ||prox(z)||^2, nnz(s) and the local slab of s for z=(w-A'lam)/tk."""
        torch = self.torch
        want = ("Axprox", "s") if want_s else ("Axprox",)
        n, m = self.n, self.m
        if self.world > 1 and hasattr(self.ops, "prox_residual_dev"):
            # the kernel leaves the norm term and the count in the tail of the message itself: no host read, no upload
            buf = torch.zeros(n + m + 2, dtype=torch.float64, device=self.w_loc.device)
            ev = self.ops.prox_residual_dev(self.w_loc, self._lam_loc(lam), self.p_loc, self.q, self.tk, self.gama, want, buf[n + m:])
            ax = ev["Axprox"]
            buf[:n] = ax[:n]
            buf[n + self.r0: n + self.r1] = ax[n:]
            self._all_reduce(buf)
            tail = buf[n + m:].tolist()                          # the one device->host read of the evaluation
            return buf[: n + m], float(tail[0]), int(round(tail[1])), ev.get("s")
        ev = self.ops.prox_residual(self.w_loc, self._lam_loc(lam), self.p_loc, self.q, self.tk, self.gama, want)
        ax = ev["Axprox"]
        if self.world == 1:
            return ax, float(ev["norm2"]), int(ev["count"]), ev.get("s")
        # ONE collective per residual: column partials (n), this rank's row sums in their place of a zero-padded
        # m-vector, and the two scalars, summed over the ranks
        buf = torch.zeros(n + m + 2, dtype=ax.dtype, device=ax.device)
        buf[:n] = ax[:n]
        buf[n + self.r0: n + self.r1] = ax[n:]
        buf[n + m:] = torch.tensor([ev["norm2"], float(ev["count"])], dtype=ax.dtype, device=ax.device)
        self._all_reduce(buf)
        tail = buf[n + m:].tolist()                              # the one device->host read of the evaluation
        return buf[: n + m], float(tail[0]), int(round(tail[1])), ev.get("s")

    def trial_batch(self, lk, zeta, delta, ll0, nt, screened=False):
        """One read-sweep of the slab for nt Armijo trials: returns (lamT, values, density) with values
        (host) = [global ||prox(z_t)||^2 (nt) ; ||lam_t||^2, wlk'lam_t pairs (2 nt)].  One all_reduce of nt
        (+1) doubles and one device->host read per batch, however many launches it takes.  screened: the
        slab goes through the screened kernels (up to 256 steps per read of the slab, gama = Inf), whose count of surviving
        entries rides in the same all_reduce and comes back as density = share of the plan's entries."""
        torch = self.torch
        lams, f0s, parts = [], [], []
        per = 256 if screened else 8
        cands = None
        for t0 in range(0, nt, per):
            k = min(per, nt - t0)
            lamT, f0 = self.ops.trial_vectors(lk, zeta, self.wlk, delta, ll0 + t0, k)
            if screened:
                out = self.ops.prox_trials_lin(self.w_loc, self._lam_loc(lk), self._lam_loc(zeta), self.p_loc, self.q, self.tk,
                                               delta, ll0 + t0, k)
                parts.append(out[:k]); cands = out[k:k + 1] if cands is None else cands + out[k:k + 1]
            else:
                lt_loc = torch.cat([lamT[:, : self.n], lamT[:, self.n + self.r0: self.n + self.r1]], dim=1).contiguous()
                parts.append(self.ops.prox_trials(self.w_loc, lt_loc, self.p_loc, self.q, self.tk, self.gama))
            lams.append(lamT); f0s.append(f0)
        if cands is not None:
            parts.append(cands)
        part = torch.cat(parts) if len(parts) > 1 else parts[0]
        self._all_reduce(part)
        vals = torch.cat([part[:nt]] + f0s).cpu().tolist()
        dens = None
        if cands is not None:
            launches = (nt + per - 1) // per
            dens = float(part[nt]) / launches / max(1.0, float(self.m) * self.n)
        return (torch.cat(lams) if len(lams) > 1 else lams[0]), vals, dens

    def assemble(self, s_loc, E=None):
        """H0 = ASAt(s,p,q) from the row-sharded active set: O(E) integers are exchanged.  ``E`` = nnz(s) over all
        ranks when the caller knows it (it rides in the residual's all_reduce): one collective instead of two."""
        lin = self.ops.active_lin(s_loc, self.m_loc, self.n, self.r0, self.m)
        if self.world == 1:
            lin_all = lin
        elif E is not None:
            lin_all = self._gather_active(lin, int(E))           # sorted: global column-major order == find(s)
        else:
            lin_all = self.torch.sort(self._all_gather_var(lin))[0]
        return self.ops.asat_from_lin(lin_all, self.p, self.q)

    def __call__(self):
        """One step from ``state['lk']`` with the random stream reset first (benchmarks, parity tests)."""
        lk_new, Fk_new, info, _ = self.step(self.lk, reset_rng=True)
        return lk_new, Fk_new, info

    def step(self, lk, pre=None, reset_rng=False, want_s_new=False):
        """One semismooth-Newton step from ``lk``.  ``pre`` = the residual evaluation at ``lk`` (the 4-tuple
        ``residual(lk, True)`` returns) when the caller already holds it -- the outer loop passes the
        evaluation that closed the previous step.  Returns ``(lk_new, Fk_new, info, post)`` with ``post`` the
        residual evaluation at ``lk_new`` (with its active set when ``want_s_new``)."""
        torch = self.torch
        bk1, tk, wlk = self.bk1, self.tk, self.wlk
        nu, delta, max_ll = 0.2, 0.9, 500
        import time as _time
        tm = {"plan": 0.0, "asat": 0.0, "amg": 0.0}
        def _lap(key, t0):
            if self.profile and torch.cuda.is_available():       # off in the timed path: a device synchronise per phase
                torch.cuda.synchronize()
            tm[key] += (_time.perf_counter() - t0) * 1e3
        if reset_rng:
            self.ops.rng_reset()
        t0 = _time.perf_counter()
        Axp, n2_old, E, s_loc = pre if pre is not None else self.residual(lk, True)  # :139-144
        Fk_old = bk1 * lk - Axp - wlk
        _lap("plan", t0); t0 = _time.perf_counter()
        H0 = self.assemble(s_loc, E)                                                 # :142
        _lap("asat", t0); t0 = _time.perf_counter()
        prob_data = {"bk1": bk1, "tk": tk, "q": self.q, "p": self.p, "T": None, "H0": H0, "z": -Fk_old}
        solve = self.ops.hybrid_amg if self.inner_solver == 4 else self.ops.hybrid_twogrid                # :161 / :178
        zeta, itamg, resamg, info = solve(prob_data, self.amg_options)                # (replicated)
        _lap("amg", t0); t0 = _time.perf_counter()
        d3 = torch.stack([lk @ lk, wlk @ lk, Fk_old @ zeta]).tolist()                # one host read for the three dots
        f0 = bk1 / 2 * d3[0] - d3[1]                                                 # :182-184
        cFk_old = f0 + 0.5 * tk * n2_old
        ress = abs(d3[2])
        # the slab pass costs 1/world of the full pass, so more trials are evaluated speculatively per
        # all_reduce / host round trip as the world grows.  gama = Inf: the screened kernels, whose candidate
        # count says how sparse the trial plans are -- 64, then 128 steps per read while under 10 % of the
        # entries survive the screen, 16 under 25 %, else the dense 8-step kernel (as ssn_linesearch does)
        screened = np.isinf(self.gama) and self.gama > 0 and hasattr(self.ops, "prox_trials_lin") and self.screen
        ll, done, passes, dens = 0, False, 0, 1.0
        # the first read evaluates ll = 0 alone (most steps accept it) -- unless the previous step's trial plans were
        # sparse (late phase: < 10 % of the entries survive the screen, long line searches): then a screened batch of
        # 64 steps costs the same one read of the slab and saves a pass and a collective
        first = 64 if (screened and self.last_density is not None and self.last_density <= 0.10) else 1
        # ... and past the ll the previous line search accepted at when that was far out (late phase: consecutive steps accept at
        # similar ll): up to 256 steps in the one read of the slab, as plan_linesearch does (csrc/plan_ops.cu)
        if first > 1 and self.last_ll is not None and self.last_ll >= 48:
            first = min(256, ((self.last_ll + 32 + 31) // 32) * 32)
        while not done:                                                              # :189-211
            lin = screened and (passes == 0 or dens <= 0.25)
            if lin and passes > 0:
                nt = (64 if passes == 1 and first == 1 else 128) if dens <= 0.10 else 16
            else:
                nt = 8 * min(self.world, 4)
            nt = min(first if passes == 0 else nt, max_ll - ll + 1)
            lamT, vals, d = self.trial_batch(lk, zeta, delta, ll, nt, screened=lin); passes += 1
            if d is not None:
                dens = d; self.last_density = d
            for t in range(nt):
                f0 = bk1 / 2 * vals[nt + 2 * t] - vals[nt + 2 * t + 1]
                if not (f0 + 0.5 * tk * vals[t] > cFk_old - nu * delta ** (ll + t) * ress) or ll + t == max_ll:
                    ll, lk_new, done = ll + t, lamT[t].clone(), True
                    break
            else:
                ll += nt
        self.last_ll = ll
        post = self.residual(lk_new, want_s_new)                                     # :212
        Fk_new = bk1 * lk_new - post[0] - wlk
        _lap("plan", t0)
        return lk_new, Fk_new, {"E": E, "ms_plan": tm["plan"], "ms_asat": tm["asat"], "ms_amg": tm["amg"], "itamg": itamg, "resamg": resamg, "info": info, "ll": ll, "ls_passes": passes,
                                "nnzH": getattr(H0, "nnz", None), "collectives": self.collectives, "zeta": zeta}, post


def make_sharded_step(state, rank, world, **kw):
    return ShardedStep(state, rank, world, **kw)
