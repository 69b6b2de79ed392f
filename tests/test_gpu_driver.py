"""GPU end-to-end parity: the device-resident APD/SsN driver against the oracle's restatement of
Class1/APD_SsN_Class1.m on the same inputs (objective <= 1e-8 relative, BASELINE.json)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,size", [("random", (40, 30)), ("grid", 8)])
def test_full_solve_matches_oracle(gpu, oracle, kind, size):
    from oracle import driver as odrv
    drv = __import__("importlib").import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    P = gpu.problems.random_problem(*size, seed=2) if kind == "random" else gpu.problems.grid_problem(size, seed=0)
    oracle.rng_reset(); gpu.rng_reset()
    ref = odrv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"])
    out = drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"])
    assert ref["stats"]["converged"] and out["stats"]["converged"]
    assert out["rel_kkt"] <= 1e-6
    f_ref, f = ref["fxk"][-1], out["fxk"][-1]
    # both runs stop at rel-KKT <= 1e-6, so their objectives agree to the solver tolerance; the
    # iterates themselves are compared over the first outer iterations, where rounding has not
    # yet been amplified by active-set changes
    assert abs(f - f_ref) <= 1e-6 * max(abs(f_ref), 1e-3)
    k = min(3, len(ref["fxk"]), len(out["fxk"]))
    assert np.allclose(out["fxk"][:k], ref["fxk"][:k], rtol=1e-8)
    assert np.allclose(out["KKT_lk"][:k], ref["KKT_lk"][:k], rtol=1e-6, atol=1e-12)
    x = out["xk"].cpu().numpy()
    A = oracle.explicit_A(P["p"], P["q"]); b = np.concatenate([P["r"], P["l"]])
    assert np.linalg.norm(A @ x - b) <= 1e-5 * (1 + np.linalg.norm(b)) and x.min() >= 0


def test_warmup_matches_oracle(gpu, oracle):
    from oracle import driver as odrv
    drv = __import__("importlib").import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    P = gpu.problems.random_problem(33, 27, seed=5)
    x_ref, l_ref = odrv.warmup_class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], 0, 100)
    x, lam = drv.warmup_class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], 0.0, 100)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    assert np.linalg.norm(lam.cpu().numpy() - l_ref) <= 1e-9 * np.linalg.norm(l_ref)
    x2, lam2 = drv.warmup_class1_unfused(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], 0.0, 100)
    assert np.linalg.norm(x2.cpu().numpy() - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    assert np.linalg.norm(lam2.cpu().numpy() - l_ref) <= 1e-9 * np.linalg.norm(l_ref)


@pytest.mark.parametrize("m,n,gama,weights", [(33, 27, np.inf, False), (130, 257, 0.05, True), (1024, 600, np.inf, True)])
def test_fused_warmup_matches_oracle(gpu, oracle, m, n, gama, weights):
    """ssn_warmup_class1 (two fused plan-wide kernels per A-ADMM iteration) against the oracle's
    line-by-line restatement of Class1/warmup_class1.m, with capacities and non-unit weights."""
    from oracle import driver as odrv
    rs = np.random.RandomState(m + n)
    P = gpu.problems.random_problem(m, n, seed=7)
    p = rs.random_sample(m) + 0.5 if weights else P["p"]
    q = rs.random_sample(n) + 0.5 if weights else P["q"]
    x_ref, l_ref = odrv.warmup_class1(P["c"], P["r"], P["l"], p, q, gama, 0, 40)
    x, lam = gpu.warmup_class1(P["c"], P["r"], P["l"], p, q, gama, 0, 40)
    assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    # lk accumulates ak/bk*(Ax(vk1) - b): differences of O(1) sums whose last bits depend on the
    # reduction order, so it is pinned to 1e-6 against the oracle and to the operator-by-operator
    # device version (same reduction kernels) much tighter
    assert np.linalg.norm(lam - l_ref) <= 1e-6 * np.linalg.norm(l_ref)
    drv = __import__("importlib").import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    x2, lam2 = drv.warmup_class1_unfused(P["c"], P["r"], P["l"], p, q, gama, 0.0, 40)
    assert np.linalg.norm(x - x2.cpu().numpy()) <= 1e-10 * np.linalg.norm(x_ref)
    assert np.linalg.norm(lam - lam2.cpu().numpy()) <= 1e-7 * np.linalg.norm(l_ref)


@pytest.mark.parametrize("kind,size,solver", [("random", (22, 18), 4), ("grid", 6, 4), ("random", (15, 12), 3), ("random", (20, 17), 5)])
def test_class2_partial_ot_solve_matches_oracle(gpu, oracle, kind, size, solver):
    """Config 3 of BASELINE.json (Class2/APD_SsN_Class2.m: partial OT through AMG4POT / PCG4POT and
    the invHHt warm start): the device driver against the oracle's restatement, same inputs."""
    from oracle import driver as odrv
    drv = __import__("importlib").import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    if kind == "grid":
        P = gpu.problems.grid_problem_pot(size, seed=0)
    else:
        rs = np.random.RandomState(4)
        m, n = size
        l = rs.random_sample(m) + 0.1; r = rs.random_sample(n) + 0.1
        P = {"c": rs.random_sample(m * n), "r": r, "l": l, "p": np.ones(m), "q": np.ones(n), "phi": np.ones(m * n),
             "mu": 0.65 * min(r.sum(), l.sum())}
    u_ref, l_ref = odrv.warmup_class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], 0, 100)
    u_w, l_w = drv.warmup_class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], 0.0, 100)
    assert np.linalg.norm(u_w.cpu().numpy() - u_ref) <= 1e-9 * np.linalg.norm(u_ref)
    assert np.linalg.norm(l_w.cpu().numpy() - l_ref) <= 1e-6 * np.linalg.norm(l_ref)
    oracle.rng_reset(); gpu.rng_reset()
    ref = odrv.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], inner_solver=solver)
    out = drv.APD_SsN_Class2(P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"], inner_solver=solver)
    assert ref["stats"]["converged"] and out["stats"]["converged"]
    assert out["rel_kkt"] <= 1e-6
    f_ref, f = ref["fxk"][-1], out["fxk"][-1]
    assert abs(f - f_ref) <= 1e-6 * max(abs(f_ref), 1e-3)
    k = min(3, len(ref["fxk"]), len(out["fxk"]))
    assert np.allclose(out["fxk"][:k], ref["fxk"][:k], rtol=1e-8)
    x = out["xk"].cpu().numpy()
    assert x.min() >= 0 and abs(P["phi"] @ x - P["mu"]) <= 1e-5 * (1 + P["mu"])      # transported mass = mu


def test_sharded_outer_loop_on_one_gpu_matches_device_driver(gpu):
    """The row-sharded Class 1 solve (sharded_driver.py, CUDA slab operators, world size 1 -- the code
    path every rank of a multi-GPU run executes, minus the collectives) against the single-GPU driver:
    same SsN step counts over the first outer iterations, objectives / KKT histories to rounding,
    and the device-generated cost slab against the host generator."""
    import importlib
    import torch
    drv = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    sd = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded_driver")
    g = 10
    P = gpu.problems.grid_problem(g, seed=0)
    m = n = g * g
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    c_slab = sd.grid_cost_slab(g, 0, m)
    assert np.allclose(c_slab.cpu().numpy(), P["c"], rtol=0, atol=3e-15)
    c_part = sd.grid_cost_slab(g, 30, 57).cpu().numpy().reshape(n, 27)
    assert np.allclose(c_part, P["c"].reshape(n, m)[:, 30:57], rtol=0, atol=3e-15)
    gpu.rng_reset()
    ref = drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], max_outer=6)
    gpu.rng_reset()
    out = sd.APD_SsN_Class1_sharded(dev(P["c"]), dev(P["r"]), dev(P["l"]), dev(P["p"]), dev(P["q"]), 0, 1, max_outer=6)
    k = len(ref["fxk"])
    assert len(out["fxk"]) == k
    assert out["stats"]["ssn_its"] == ref["stats"]["ssn_its"]
    assert np.allclose(out["fxk"], ref["fxk"], rtol=1e-9, atol=1e-12)
    assert np.allclose(out["KKT_lk"], ref["KKT_lk"], rtol=1e-6, atol=1e-11)
    assert np.allclose(out["KKT_xk"], ref["KKT_xk"], rtol=1e-6, atol=1e-11)
    assert np.allclose(out["lk"].cpu().numpy(), ref["lk"].cpu().numpy(), rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("m,n,gama", [(100, 100, np.inf), (77, 53, np.inf), (64, 48, 0.02)])
def test_staged_warm_start_equals_the_fused_warm_start(gpu, m, n, gama):
    """ssn_warm_stage (one fused stage per call, the loop and the (n+m)-sized updates on the host side, as the
    row-sharded driver runs it) against ssn_warmup_class1 (the whole loop in the library): same kernels, same
    scalars => the same iterates up to the rounding of invAAt / the dual update, which the two paths evaluate
    with different small kernels."""
    import importlib
    import torch
    sd = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded_driver")
    sharded = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.sharded")
    P = gpu.problems.random_problem(m, n, seed=5)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    c, r, l, p, q = (dev(P[k]) for k in ("c", "r", "l", "p", "q"))
    x_ref, lk_ref = gpu.warmup_class1(c, r, l, p, q, gama, 0.0, 25)
    A = sd.SlabAlgebra(0, 1, p, q, sharded._CudaOps(), None, torch)
    x, lk = sd.warmup_class1_sharded_fused(A, c, torch.cat([r, l]), gama, 25)
    assert torch.allclose(x, x_ref, rtol=1e-10, atol=1e-13)
    assert torch.allclose(lk, lk_ref, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("inner_solver,precd", [(2, 2), (2, 4), (3, 2), (5, None)])
def test_every_selectable_inner_solver_reaches_the_same_optimum(gpu, inner_solver, precd):
    """Class1/APD_SsN_Class1.m:66-70: inner_solver 2 (PCG on Jk, here with the Jacobi and the ichol
    preconditioner), 3 (aug_PCG) and 5 (Hybrid_twogrid) against the default 4 (Hybrid_AMG) on a small grid
    problem: all solve the same Newton systems to 1e-11, so the solves converge to the same objective."""
    import importlib
    drv = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    P = gpu.problems.grid_problem(8, seed=0)
    gpu.rng_reset()
    ref = drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"])
    gpu.rng_reset()
    po = None if precd is None else {"retol": 1e-11, "maxit": 10000, "precd": precd, "guess": None}
    out = drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], inner_solver=inner_solver, pcg_options=po)
    assert ref["stats"]["converged"] and out["stats"]["converged"]
    assert abs(out["fxk"][-1] - ref["fxk"][-1]) <= 1e-6 * max(abs(ref["fxk"][-1]), 1e-3)
    k = min(3, len(ref["fxk"]), len(out["fxk"]))
    assert np.allclose(out["fxk"][:k], ref["fxk"][:k], rtol=1e-7)


def test_one_call_ssn_step_equals_the_operator_level_step(gpu):
    """ssn_ssn_step_class1 (one library call: residual -> ASAt -> Hybrid_AMG -> Armijo loop -> residual) and its
    host-buffer variant against driver.ssn_step, which makes the same step out of the operator-level calls: same discrete
    facts, same duals."""
    import importlib
    import torch
    drv = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    P = gpu.problems.grid_problem(24, seed=1)
    gpu.rng_reset()
    st = drv.capture_state(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], outer=3, ssn_it=1)
    gpu.rng_reset(); lk_a, Fk_a, ia = drv.ssn_step(st)
    gpu.rng_reset(); lk_b, Fk_b, ib = gpu.ssn_step_class1(st["wk"], st["lk"], st["wlk"], st["p"], st["q"], st["bk1"], st["tk"])
    gpu.rng_reset(); lk_c, Fk_c, ic = gpu.ssn_step_class1(st["wk"].cpu(), st["lk"].cpu(), st["wlk"].cpu(), st["p"].cpu(), st["q"].cpu(),
                                                          st["bk1"], st["tk"], host_call=True)
    for info in (ib, ic):
        assert (info["E"], info["itamg"], info["ll"], info["info"][0]) == (int(ia["E"]), int(ia["itamg"]), int(ia["ll"]), int(ia["info"][0]))
    scale = float(lk_a.abs().max())
    assert float((lk_b - lk_a).abs().max()) <= 1e-13 * scale and float((lk_c.cuda() - lk_a).abs().max()) <= 1e-13 * scale
    assert float((Fk_b - Fk_a).abs().max()) <= 1e-10 * float(Fk_a.abs().max()) + 1e-14


@pytest.mark.parametrize("m,n", [(64, 48), (45, 70)])
def test_class2_one_call_entry_points_match_the_operator_loop(gpu, m, n):
    """ssn_warmup_class2 / ssn_apd_ssn_class2 / ssn_ssn_step_class2 (the Class 2 script, its warm start and one SsN step as
    single library calls over the fused PHI kernels) against the same script as a Python loop over the operators and the
    operator-by-operator warm start (driver.APD_SsN_Class2_loop, warmup_class2_unfused, ssn_step_class2_ops)."""
    import torch
    drv = __import__("importlib").import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    rs = np.random.RandomState(m + n)
    l = rs.random_sample(m) + 0.1; r = rs.random_sample(n) + 0.1
    P = {"c": rs.random_sample(m * n), "r": r, "l": l, "p": np.ones(m), "q": np.ones(n), "phi": np.ones(m * n),
         "mu": 0.6 * min(r.sum(), l.sum())}
    a = (P["c"], P["r"], P["l"], P["p"], P["q"], P["mu"], P["phi"])
    u1, l1 = drv.warmup_class2(*a, 0.0, 60)
    u2, l2 = drv.warmup_class2_unfused(*a, 0.0, 60)
    assert float(torch.linalg.norm(u1 - u2)) <= 1e-9 * float(torch.linalg.norm(u2))
    assert float(torch.linalg.norm(l1 - l2)) <= 1e-7 * float(torch.linalg.norm(l2))
    # the loop from the trivial start (the warm start's closed form is ill conditioned, tests/test_gpu_traces.py)
    gpu.rng_reset()
    one = gpu.APD_SsN_Class2(*a, warm_maxit=0)
    gpu.rng_reset()
    seen = []
    loop = drv.APD_SsN_Class2_loop(*a, warm_maxit=0, on_ssn_step=lambda st: seen.append(st["k"]))
    assert one["stats"]["converged"] and loop["stats"]["converged"] and abs(one["outer_its"] - loop["outer_its"]) <= 1
    # the two callers sum their dot products in different orders: decisions are pinned while |F| is far above SsN_Tol1
    K = 20
    i1, i2 = one["stats"]["ssn_its"], loop["stats"]["ssn_its"]
    L = min(len(i1), len(i2))
    assert i1[:K] == i2[:K] and max(abs(a_ - b_) for a_, b_ in zip(i1[:L], i2[:L])) <= 3 and len(seen) == sum(i2)   # late steps: |F| at the rounding level
    assert np.allclose(one["fxk"][:K], loop["fxk"][:K], rtol=1e-9, atol=1e-12)
    assert abs(one["fxk"][-1] - loop["fxk"][-1]) <= 1e-6 * max(abs(loop["fxk"][-1]), 1e-3)
    K1, K2 = np.array(one["KKT"][:K]), np.array(loop["KKT"][:K])
    assert np.allclose(K1, K2, rtol=1e-6, atol=1e-10)
    ns = sum(i2[:K])
    assert [s_[:6] for s_ in one["stats"]["steps"][:ns]] == [tuple(int(v) for v in s_[:6]) for s_ in loop["stats"]["steps"][:ns]]
    assert np.allclose(one["lk"].cpu().numpy(), loop["lk"].cpu().numpy(), rtol=1e-4, atol=1e-6)
    host = gpu.APD_SsN_Class2(*a, warm_maxit=0, host_call=True, max_outer=3)
    assert np.allclose(host["fxk"], one["fxk"][:4], rtol=1e-12) and isinstance(host["uk"], np.ndarray)
    # one SsN step at the trivial state
    Pt = {k: torch.from_numpy(np.asarray(v, dtype=np.float64)).cuda() if not np.isscalar(v) else v for k, v in P.items()}
    st = drv.class2_trivial_state(Pt)
    gpu.rng_reset(); lk_a, Fk_a, ia = drv.ssn_step_class2(st)
    gpu.rng_reset(); lk_b, Fk_b, ib = drv.ssn_step_class2_ops(st)
    assert ia["E"] == ib["E"] and ia["ll"] == ib["ll"] and ia["itamg"] == ib["itamg"] and ia["nnzH"] == ib["nnzH"]
    assert torch.allclose(lk_a, lk_b, rtol=1e-11, atol=1e-13) and torch.allclose(Fk_a, Fk_b, rtol=1e-9, atol=1e-12)
    hst = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
    gpu.rng_reset(); lk_h, Fk_h, ih = drv.ssn_step_class2(hst, host_call=True)
    assert not lk_h.is_cuda and torch.equal(lk_h, lk_a.cpu()) and torch.equal(Fk_h, Fk_a.cpu())
