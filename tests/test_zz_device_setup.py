"""Device-side setup paths of the non-default solvers: SSOR / IC(0) factors and their dependency levels built on the
device (csrc/trifactor.cu; the library's default since round 2, ``ssn_set_device_setup(0)`` / SSN_DEVICE_SETUP=0 selects
the host construction that is kept as the cross-check) and the assembled KKT matrix ``ssn_jk_system``.  These tests run
them against the host construction (bit for bit) and the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _systems(oracle):
    g = 24
    T = sp.diags([-np.ones(g - 1), 2 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1])
    A = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g)) + 0.05 * sp.identity(g * g)).tocsc()
    rs = np.random.RandomState(6)
    m, n = 130, 110
    s = rs.random_sample(m * n) < 0.05
    H0 = oracle.ASAt(s, rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5)
    Jk = (0.3 * sp.identity(m + n) + H0 / 0.8).tocsc()
    return A, Jk, rs


@pytest.mark.parametrize("precd", [3, 4])
def test_device_factors_equal_host_factors(gpu, oracle, precd):
    """Same arithmetic entry by entry => the PCG run is identical whether the factors come from the host or the device."""
    A, Jk, rs = _systems(oracle)
    try:
        for M in (A, Jk):
            b = rs.standard_normal(M.shape[0])
            o = {"retol": 1e-11, "maxit": 2000, "precd": precd, "guess": None}
            gpu.set_device_setup(False)
            d0, it0, res0, resk0 = gpu.PCG(M, b, o)
            gpu.set_device_setup(True)
            d1, it1, res1, resk1 = gpu.PCG(M, b, o)
            assert it1 == it0 and res1 == res0
            assert np.array_equal(np.asarray(resk1), np.asarray(resk0))
            assert np.array_equal(np.asarray(d1), np.asarray(d0))
            d_ref, it_ref, _, _ = oracle.PCG(M, b, o)
            assert abs(it1 - it_ref) <= 1
            assert np.linalg.norm(d1 - d_ref) <= 1e-8 * np.linalg.norm(d_ref)
        if precd == 4:
            gpu.set_device_setup(True)
            with pytest.raises(Exception) as e:
                gpu.PCG((A - 10 * sp.identity(A.shape[0])).tocsc(), rs.standard_normal(A.shape[0]),
                        {"retol": 1e-11, "maxit": 10, "precd": 4, "guess": None})
            assert "SSN_E_NOT_SPD" in str(e.value)
    finally:
        gpu.set_device_setup(True)


def test_device_factors_deep_dependency_chain(gpu, oracle):
    """A tridiagonal matrix: n dependency levels (the level relaxation needs n sweeps), one row per level."""
    n = 300
    A = sp.diags([-np.ones(n - 1), 2.5 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsc()
    b = np.random.RandomState(2).standard_normal(n)
    try:
        for precd in (3, 4):
            o = {"retol": 1e-12, "maxit": 500, "precd": precd, "guess": None}
            gpu.set_device_setup(False)
            d0, it0, _, _ = gpu.PCG(A, b, o)
            gpu.set_device_setup(True)
            d1, it1, _, _ = gpu.PCG(A, b, o)
            assert it1 == it0 and np.array_equal(np.asarray(d1), np.asarray(d0))
            if precd == 4:
                assert it1 <= 3                                       # IC(0) of a tridiagonal matrix is its exact Cholesky factor
    finally:
        gpu.set_device_setup(True)


@pytest.mark.parametrize("with_T,drop_diag", [(False, False), (True, False), (True, True)])
def test_jk_system(gpu, oracle, with_T, drop_diag):
    """ssn_jk_system against bk1*speye + (T+H0)/tk formed with SciPy: pattern exact, values bit for bit; a node without
    any active entry has no diagonal in H0 but one (bk1 + t/tk) in Jk."""
    rs = np.random.RandomState(11)
    m, n = 90, 70
    S = rs.random_sample((m, n)) < 0.06
    if drop_diag:
        S[5, :] = False; S[:, 9] = False                              # an isolated row node and an isolated column node
    p, q = rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5
    H0 = oracle.ASAt(S.reshape(-1, order="F"), p, q)
    bk1, tk = 0.37, 0.81
    t = rs.random_sample(m + n) if with_T else None
    pd = {"bk1": bk1, "tk": tk, "p": p, "q": q, "T": t, "H0": H0, "z": np.zeros(m + n)}
    Jk = gpu.jk_system(pd).to_scipy().tocsr(); Jk.sort_indices()
    Hc = H0.tocsr(); Hc.sort_indices()
    TH = Hc if t is None else (sp.diags(t).tocsr() + Hc)
    ref = (bk1 * sp.identity(m + n, format="csr") + TH / tk).tocsr(); ref.sort_indices()
    assert np.array_equal(Jk.indptr, ref.indptr) and np.array_equal(Jk.indices, ref.indices)
    assert np.array_equal(Jk.data, ref.data)
    if drop_diag:
        assert Hc[n + 5, n + 5] == 0 and Jk[n + 5, n + 5] == bk1 + (0.0 if t is None else t[n + 5]) / tk


def test_driver_inner_solver_2_runs_pcg_on_the_device_assembled_jk(gpu, oracle):
    """The Class 1 driver with inner_solver = 2 (Class1/APD_SsN_Class1.m:149-152) assembles Jk with ssn_jk_system: its
    first Newton direction equals PCG on bk1*speye + (T+H0)/tk formed with SciPy from the same state."""
    import importlib
    drv = importlib.import_module("codes-of-ipd-ssn-amg-method_b200.driver")
    P = gpu.problems.grid_problem(8, seed=0)
    po = {"retol": 1e-11, "maxit": 10000, "precd": 2, "guess": None}
    seen = []

    def hook(st):
        if not seen:
            seen.append({"H0": st["H0"].to_scipy().tocsr(), "Fk": st["Fk"].cpu().numpy().copy(), "bk1": st["bk1"], "tk": st["tk"]})
    gpu.rng_reset()
    out = drv.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], inner_solver=2, pcg_options=po, max_outer=2,
                             on_ssn_step=hook)
    st = seen[0]
    N = st["H0"].shape[0]
    Jk = (st["bk1"] * sp.identity(N, format="csr") + st["H0"] / st["tk"]).tocsc()
    d_ref, it_ref, _, _ = oracle.PCG(Jk, -st["Fk"], po)
    d, it, _, _ = gpu.PCG(Jk, -st["Fk"], po)
    assert abs(it - it_ref) <= 1 and np.linalg.norm(np.asarray(d) - d_ref) <= 1e-8 * np.linalg.norm(d_ref)
    assert out["stats"]["lin_its"][0][0] == it
