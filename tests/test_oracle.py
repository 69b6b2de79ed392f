"""CPU tests of the oracle itself (no GPU): the identities the reference's own comments state
(SURVEY.md section 4), its frozen conventions, and the committed golden vectors."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import GOLDEN, golden_systems, load_system, random_active_problem

AMG_OPTS = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 0.25, "smoth": 5, "cycle": "w", "isnsp": 1,
            "inter": 1, "guess": None}


def test_matlab_stream_first_values(oracle):
    oracle.rng_reset()
    v = oracle.rand(5)
    # MATLAB's well-known first draws of `rand` after start-up (mt19937ar, seed 0)
    assert np.allclose(v, [0.8147236863931789, 0.9057919370756192, 0.12698681629350606, 0.9133758561390194,
                           0.6323592462254095], rtol=0, atol=1e-15)
    oracle.rng_reset()
    assert np.array_equal(oracle.rand(2000), np.random.RandomState(5489).random_sample(2000))
    first = np.load(os.path.join(GOLDEN, "matlab_rand_first.npz"))["first"]
    oracle.rng_reset()
    assert np.array_equal(oracle.rand(1000), first)


def test_ax_aty_equal_explicit_matrix_and_adjoint(oracle):
    rs = np.random.RandomState(0)
    m, n = 13, 9
    p, q = rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5
    A = oracle.explicit_A(p, q)
    x, y = rs.standard_normal(m * n), rs.standard_normal(m + n)
    assert np.allclose(oracle.Ax(x, p, q), A @ x, rtol=1e-13)
    assert np.allclose(oracle.Aty(y, p, q), A.T @ y, rtol=1e-13)
    assert abs(oracle.Ax(x, p, q) @ y - x @ oracle.Aty(y, p, q)) < 1e-11
    assert np.allclose(oracle.Ax(sp.csc_matrix((m * n, 1)), p, q), 0.0)      # warmup_class1.m:29


def test_c_plan_kernels_match_numpy(oracle):
    from oracle import _ck
    rs = np.random.RandomState(1)
    m, n = 37, 21
    p, q = rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5
    x, y = rs.standard_normal(m * n), rs.standard_normal(m + n)
    out = np.empty(m + n); _ck.lib().orc_ax(x, p, q, m, n, out)
    assert np.allclose(out, oracle.Ax(x, p, q), rtol=1e-13)
    z = np.empty(m * n); _ck.lib().orc_aty(y, p, q, m, n, z)
    assert np.array_equal(z, oracle.Aty(y, p, q))


@pytest.mark.parametrize("weights", [False, True])
def test_asat_is_A_S_At(oracle, weights):
    m, n = 17, 23
    s, p, q = random_active_problem(m, n, 0.15, 3, weights)
    H = oracle.ASAt(s, p, q)
    A = oracle.explicit_A(p, q)
    E = (A @ sp.diags(s.astype(float)) @ A.T).tocsc(); E.eliminate_zeros(); E.sort_indices()
    assert np.array_equal(H.indptr, E.indptr) and np.array_equal(H.indices, E.indices)       # pattern exact
    assert np.allclose(H.data, E.data, rtol=1e-10)
    z = np.random.RandomState(0).standard_normal(m + n)
    assert H.nnz == 2 * int(s.sum()) + np.count_nonzero(H.diagonal())


def test_asatz_identity_and_bug(oracle):
    m = n = 12
    s, p, q = random_active_problem(m, n, 0.2, 4, True)
    z = np.random.RandomState(1).standard_normal(m + n)
    assert np.allclose(oracle.ASAtz(z, s, p, p), oracle.ASAt(s, p, p) @ z, rtol=1e-12)       # p == q: harmless
    assert not np.allclose(oracle.ASAtz(z, s, p, q), oracle.ASAt(s, p, q) @ z)               # ASAtz.m:21 typo
    with pytest.raises(ValueError):
        oracle.ASAtz(np.zeros(7), np.zeros(12, bool), np.ones(3), np.ones(4))


def test_closed_form_inverses(oracle):
    rs = np.random.RandomState(2)
    m, n = 11, 8
    p, q = rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5
    A = oracle.explicit_A(p, q)
    x = rs.standard_normal(m + n)
    M = sp.diags(np.concatenate([0.7 * np.ones(n), 1.3 * np.ones(m)])) + A @ A.T              # invAAt.m:2
    assert np.allclose(M @ oracle.invAAt(x, p, q, 0.7, 1.3), x, rtol=1e-10)
    phi = rs.random_sample(m * n); sg = 0.6
    G = sp.vstack([A, sp.csr_matrix(phi[None, :])])
    IY = sp.vstack([sp.identity(n), sp.csr_matrix((m + 1, n))]); IZ = sp.vstack([sp.csr_matrix((n, m)), sp.identity(m), sp.csr_matrix((1, m))])
    Hm = sp.hstack([G, IY, IZ])
    v = rs.standard_normal(m + n + 1)
    assert np.allclose((sg * sp.identity(m + n + 1) + Hm @ Hm.T) @ oracle.invHHt(v, p, q, sg, phi), v, rtol=1e-9)


def test_spgemm_order_matches_scipy_and_drops_zeros(oracle):
    from oracle.amg import spgemm
    rs = np.random.RandomState(5)
    A = sp.random(60, 80, density=0.1, random_state=rs, format="csc"); B = sp.random(80, 70, density=0.1, random_state=rs, format="csc")
    C = spgemm(A, B); R = (A @ B).tocsc(); R.eliminate_zeros(); R.sort_indices()
    assert np.array_equal(C.indptr, R.indptr) and np.array_equal(C.indices, R.indices)
    assert np.array_equal(C.data, R.data)       # SciPy's csc*csc walks k in the same ascending order
    X = sp.csc_matrix(np.array([[1.0, -1.0], [2.0, 3.0]])); Y = sp.csc_matrix(np.array([[1.0, 5.0], [1.0, 7.0]]))
    assert spgemm(X, Y).nnz == 3


def _system(oracle, m, n, density, seed, bk1=0.05, tk=0.8, connect=True):
    s, p, q = random_active_problem(m, n, density, seed)
    if not connect:
        S = s.reshape((m, n), order="F").copy()
        S[: m // 2, n // 2:] = False; S[m // 2:, : n // 2] = False; S[0, :] = False; S[:, 0] = False; S[0, 0] = True
        s = S.reshape(-1, order="F")
    H0 = oracle.ASAt(s, p, q)
    return {"bk1": bk1, "tk": tk, "p": p, "q": q, "T": sp.diags(np.zeros(m + n)), "H0": H0,
            "z": np.random.RandomState(seed).standard_normal(m + n), "s": s}


def test_rescaled_laplacian_properties(oracle):
    from oracle.solvers import rescaled_system
    pd = _system(oracle, 40, 30, 0.1, 1)
    qp, A0, Qd, Kd, Ae, f = rescaled_system(pd)
    assert np.allclose(A0 @ np.ones(70), 0.0, atol=1e-12)          # A0 is a graph Laplacian
    blocks, sizes, p, r = oracle.components(A0)
    Y = sp.csc_matrix((np.ones(70), (np.arange(70), blocks - 1)))
    assert abs(A0 @ Y).max() < 1e-12                               # aug_PCG.m:27 "A0*Y = 0"
    assert sorted(p.tolist()) == list(range(70)) and r[-1] == 70
    for k in range(len(sizes)):
        mem = p[r[k]:r[k + 1]]
        assert np.all(np.diff(mem) > 0) and (k == 0 or mem[0] > p[r[k - 1]])


@pytest.mark.parametrize("connect", [True, False])
def test_hybrid_amg_solves_the_newton_system(oracle, connect):
    m, n = 260, 240
    pd = _system(oracle, m, n, 0.02, 7, connect=connect)
    oracle.rng_reset()
    zeta, it, res, info = oracle.Hybrid_AMG(pd, AMG_OPTS)
    Jk = pd["bk1"] * sp.identity(m + n) + pd["H0"] / pd["tk"]
    assert np.linalg.norm(Jk @ zeta - pd["z"]) <= 1e-9 * np.linalg.norm(pd["z"])
    assert (info[0] == 1) == connect and 0 < it < 30
    z2, it2, res2, info2 = oracle.aug_PCG(pd, {"retol": 1e-11, "maxit": 1e4, "precd": 2, "guess": None})
    assert np.linalg.norm(z2 - zeta) <= 1e-6 * np.linalg.norm(zeta)


def test_pcg_variants(oracle):
    from oracle.solvers import rescaled_system
    pd = _system(oracle, 80, 70, 0.05, 3, bk1=0.3)
    _, _, _, _, Ae, f = rescaled_system(pd)
    x = spla.spsolve(Ae.tocsc(), f)
    for precd, extra in [(1, {}), (2, {}), (3, {}), (5, {"nf": 70})]:
        d, it, res, resk = oracle.PCG(Ae, f, dict({"retol": 1e-11, "maxit": 1e4, "precd": precd, "guess": None}, **extra))
        assert np.linalg.norm(d - x) <= 1e-7 * np.linalg.norm(x), precd
        assert res <= 1e-11 and resk[it - 1] == res
    d, it, res, _ = oracle.PCG(Ae, np.zeros(150))
    assert it == 0 and np.isnan(res)                               # PCG.m:87: 0/0
    with pytest.raises(ValueError):
        oracle.PCG(Ae, f, {"retol": None, "maxit": None, "precd": 5, "guess": None})


def test_cube_root_threshold_quirk(oracle):
    from oracle.amg import coarsest_size_threshold
    assert coarsest_size_threshold(32768) == 32                    # glibc pow(32768,1/3) = 31.999999999999996
    assert coarsest_size_threshold(1000) == 10 and coarsest_size_threshold(999) == 10
    assert coarsest_size_threshold(27) == 4


def test_pot_bordering(oracle):
    m, n = 60, 50
    pd = _system(oracle, m, n, 0.06, 9, bk1=0.2)
    pd["T"] = sp.diags((np.random.RandomState(0).random_sample(m + n) > 0.5).astype(float))
    pd["phi"] = np.random.RandomState(5).random_sample(m * n) + 0.5
    pd["z"] = np.random.RandomState(6).standard_normal(m + n + 1)
    oracle.rng_reset()
    zeta, it, res, info = oracle.AMG4POT(pd, dict(AMG_OPTS, maxit=40, smoth=10), "amg")
    s = pd["s"].astype(float); A = oracle.explicit_A(pd["p"], pd["q"]); ss = A @ (s * pd["phi"])
    cH = sp.bmat([[pd["T"] + pd["H0"], ss[:, None]], [ss[None, :], [[pd["phi"] @ (s * pd["phi"])]]]])
    He = pd["bk1"] * sp.identity(m + n + 1) + cH / pd["tk"]
    assert np.linalg.norm(He @ zeta - pd["z"]) <= 1e-8 * np.linalg.norm(pd["z"])
    z2, *_ = oracle.PCG4POT(pd, {"retol": 1e-11, "maxit": 1e4, "precd": 2, "guess": None})
    assert np.linalg.norm(z2 - zeta) <= 1e-6 * np.linalg.norm(zeta)


@pytest.mark.parametrize("path", golden_systems("grid12") + golden_systems("bundled500")[:3])
def test_oracle_reproduces_golden_systems(oracle, path):
    d = load_system(path)
    m, n = d["m"], d["n"]
    H0 = oracle.ASAt(d["s"], np.ones(m), np.ones(n))
    assert np.array_equal(H0.indptr, d["H_indptr"]) and np.array_equal(H0.indices, d["H_indices"])
    pd = {"bk1": float(d["bk1"]), "tk": float(d["tk"]), "p": np.ones(m), "q": np.ones(n),
          "T": sp.diags(np.zeros(m + n)), "H0": H0, "z": d["z"]}
    oracle.rng_reset()
    zeta, it, res, info = oracle.Hybrid_AMG(pd, AMG_OPTS)
    assert it == int(d["it"]) and list(info) == list(d["info"])
    assert np.array_equal(zeta, d["zeta"])                          # the oracle is deterministic
    assert oracle.GLOBAL_STREAM.drawn == int(d["drawn"])


def test_full_solve_small_lp_against_highs(oracle):
    """End-to-end pin: the restated APD/SsN/AMG loop reaches the LP optimum an independent solver
    (SciPy HiGHS) finds -- the check Class1/APD_SsN_Class1.m:42-50 itself suggests."""
    from scipy.optimize import linprog
    from oracle import driver
    import ssnamg
    P = ssnamg.problems.random_problem(14, 11, seed=1)
    oracle.rng_reset()
    out = driver.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"])
    assert out["stats"]["converged"]
    A = oracle.explicit_A(P["p"], P["q"]); b = np.concatenate([P["r"], P["l"]])
    lp = linprog(P["c"], A_eq=A.toarray(), b_eq=b, bounds=(0, None), method="highs")
    assert abs(out["fxk"][-1] - lp.fun) <= 1e-5 * max(abs(lp.fun), 1.0)


def test_bundled_solve_summary_fixture():
    d = np.load(os.path.join(GOLDEN, "bundled500_summary.npz"))
    assert int(d["outer_its"]) == 58 and float(d["rel_kkt"]) <= 1e-6
    assert abs(float(d["f"]) - 1.1260464956) < 1e-9 and int(d["nnz"]) == 999       # basic solution, m+n-1


def test_bundled_class2_summary_fixture():
    """The oracle's Class2 solve on the reference's bundled Class2/InputData/data4-500.mat (generated by
    tests/golden/make_golden.py class2): converges like the reference claims for its own example."""
    d = np.load(os.path.join(GOLDEN, "bundled500_class2_summary.npz"))
    assert int(d["outer_its"]) == 53 and float(d["rel_kkt"]) <= 1e-6
    assert abs(float(d["fxk"][-1]) - 0.25633496670) < 1e-9
    assert abs(float(d["mass"]) - float(d["mu"])) <= 1e-5 * float(d["mu"])      # transported mass = mu = 161.2933


def test_class2_partial_ot_oracle_against_highs(oracle):
    """The oracle's restatement of Class2/APD_SsN_Class2.m (AMG4POT inner solves, invHHt warm start)
    reaches the optimum of the partial-OT LP  min c'x  s.t. Ax + [y;z] = [r;l], phi'x = mu, x,y,z >= 0
    found by an independent solver (SciPy HiGHS)."""
    import scipy.sparse as sp
    from scipy.optimize import linprog
    from oracle import driver as odrv
    rs = np.random.RandomState(3)
    m, n = 14, 11; N = m + n
    c = rs.random_sample(m * n); l = rs.random_sample(m) + 0.1; r = rs.random_sample(n) + 0.1
    phi = np.ones(m * n); mu = 0.65 * min(r.sum(), l.sum())
    oracle.rng_reset()
    out = odrv.APD_SsN_Class2(c, r, l, np.ones(m), np.ones(n), mu, phi)
    assert out["stats"]["converged"] and out["rel_kkt"] <= 1e-6
    A = oracle.explicit_A(np.ones(m), np.ones(n))
    Aeq = sp.bmat([[A, sp.identity(N)], [sp.csr_matrix(phi[None, :]), sp.csr_matrix((1, N))]]).tocsr()
    res = linprog(np.concatenate([c, np.zeros(N)]), A_eq=Aeq, b_eq=np.concatenate([r, l, [mu]]), bounds=(0, None), method="highs")
    assert res.status == 0
    assert abs(out["fxk"][-1] - res.fun) <= 1e-6 * max(1.0, abs(res.fun))
    assert abs(phi @ out["xk"] - mu) <= 1e-6 * (1 + mu)



def test_twogrid_bigph_and_hybrid_twogrid_solve_the_kkt_system():
    """AMG/twogrid_bigph.m and Hybrid_twogrid.m (inner_solver = 5) in the oracle: the two-level method solves
    the rescaled KKT system of a connected and of a disconnected active set to its tolerance, returns the
    same solution as Hybrid_AMG, and honours the function's own option defaults."""
    import scipy.sparse as sp
    import oracle
    rs = np.random.RandomState(4)
    m, n = 60, 50
    opts = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 0.25, "smoth": 5, "cycle": "w", "isnsp": 1, "inter": 1, "guess": None}
    for dens in (0.25, 0.02):
        s = rs.random_sample(m * n) < dens
        p = rs.random_sample(m) + 0.5; q = rs.random_sample(n) + 0.5
        H0 = oracle.ASAt(s, p, q)
        z = rs.standard_normal(m + n)
        pd = {"bk1": 0.3, "tk": 0.8, "p": p, "q": q, "T": sp.diags(np.zeros(m + n)), "H0": H0, "z": z}
        oracle.rng_reset()
        zeta, it, res, info = oracle.Hybrid_twogrid(pd, opts)
        Jk = 0.3 * sp.identity(m + n) + H0 / 0.8
        assert np.linalg.norm(Jk @ zeta - z) <= 1e-8 * np.linalg.norm(z)
        oracle.rng_reset()
        zeta2, _, _, info2 = oracle.Hybrid_AMG(pd, opts)
        assert np.array_equal(info, info2)
        assert np.allclose(zeta, zeta2, rtol=1e-7, atol=1e-10)
    # twogrid_bigph on its own: a bigraph system [V -U; -U' T] with the first Nf nodes as F nodes
    s = rs.random_sample(m * n) < 0.3
    H0 = oracle.ASAt(s, np.ones(m), np.ones(n))
    A = (0.1 * sp.identity(m + n) + H0.multiply(1.0) ).tocsc()
    A = (sp.diags(A.diagonal()) - (A - sp.diags(A.diagonal()))).tocsc()          # Laplacian sign pattern
    b = rs.standard_normal(m + n)
    x, it, rel, relk, rhok = oracle.twogrid_bigph(A, b, {"retol": 1e-10, "maxit": 40, "fnode": n, "smoth": 3, "isnsp": 0, "guess": None})
    assert rel <= 1e-10 and len(relk) == it + 1 and np.linalg.norm(A @ x - b) <= 1e-9 * np.linalg.norm(b)
    x2, it2, rel2, _, _ = oracle.twogrid_bigph(A, b, {"retol": None, "maxit": 3, "fnode": n, "smoth": None, "isnsp": None, "guess": None})
    assert it2 == 3 and rel2 > 0                              # retol [] -> 0: runs to maxit


def test_generic_twogrid_on_a_grid_laplacian():
    """AMG/twogrid.m with bigph = 0 in the oracle: damped Jacobi + MIS coarsening on a (nearly singular,
    hence isnsp = 1) shifted 2-D grid Laplacian converges; bigph = 1 without fnode is the reference's error."""
    import scipy.sparse as sp
    import oracle
    from oracle.amg import AMGError
    g = 14
    T = sp.diags([-np.ones(g - 1), 2 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1])
    A = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g)) + 1e-3 * sp.identity(g * g)).tocsc()
    b = np.random.RandomState(1).standard_normal(g * g)
    oracle.rng_reset()
    x, it, rel, relk, rhok = oracle.twogrid(A, b, {"retol": 1e-10, "bigph": 0, "maxit": 60, "smoth": 3, "isnsp": 1, "guess": None})
    assert rel <= 1e-10 and it < 60 and np.linalg.norm(A @ x - b) <= 1e-9 * np.linalg.norm(b)
    with pytest.raises(AMGError):
        oracle.twogrid(A, b, {"retol": 1e-10, "bigph": 1, "maxit": 5, "smoth": 3, "isnsp": 0, "guess": None})


def test_pcg_preconditioners_3_and_4():
    """PCG.m precd 3 (SSOR, :39-44, 96-99) and 4 (ichol, :45-51, 100-101) in the oracle: the IC(0) factor
    reproduces H on H's pattern, both preconditioners cut the iteration count of plain CG and solve the
    system; a matrix that is not positive definite is ichol's 'nonpositive pivot' error."""
    import scipy.sparse as sp
    import oracle
    from oracle.pcg import ichol0
    g = 12
    T = sp.diags([-np.ones(g - 1), 2 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1])
    A = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g)) + 0.05 * sp.identity(g * g)).tocsc()
    L = ichol0(A)
    assert (sp.triu(L, 1)).nnz == 0 and L.nnz == sp.tril(A).nnz
    R = (L @ L.T - A).tocsr(); pat = A.tocsr()
    assert max(abs(R[i, j]) for i, j in zip(*pat.nonzero())) <= 1e-13
    b = np.random.RandomState(0).standard_normal(g * g)
    its = {}
    for precd in (1, 3, 4):
        d, it, res, resk = oracle.PCG(A, b, {"retol": 1e-11, "maxit": 1000, "precd": precd, "guess": None})
        assert np.linalg.norm(A @ d - b) <= 1e-9 * np.linalg.norm(b)
        its[precd] = it
    assert its[3] < its[1] and its[4] < its[1]
    with pytest.raises(ValueError):
        ichol0((A - 10 * sp.identity(g * g)).tocsc())
