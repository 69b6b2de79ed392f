"""GPU parity tests, plan operators: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs.  Tolerances: bit-exact for logical / integer / pattern results and for every
value whose rounding sequence is fixed (Aty, z, prox, ASAt); <= 1e-10 relative for reductions
(BASELINE.json north_star: "<=1e-10 on operators")."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_active_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-10
SHAPES = [(1, 1), (2, 3), (7, 5), (128, 64), (250, 130), (500, 500), (1027, 517), (2048, 300), (96, 2500)]


def close(a, b, rtol=RTOL):
    a = np.asarray(a); b = np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return np.all(np.abs(a - b) <= rtol * scale)


def weights(m, n, seed, unit):
    rs = np.random.RandomState(seed)
    return (np.ones(m), np.ones(n)) if unit else (rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5)


@pytest.mark.parametrize("m,n", SHAPES)
@pytest.mark.parametrize("unit", [True, False])
def test_ax_matches_oracle(gpu, oracle, m, n, unit):
    rs = np.random.RandomState(m * 131 + n)
    x = rs.standard_normal(m * n); p, q = weights(m, n, 1, unit)
    y = gpu.Ax(x, p, q)
    assert y.shape == (n + m,)
    assert close(y, oracle.Ax(x, p, q))


@pytest.mark.parametrize("m,n", SHAPES)
def test_aty_bit_exact(gpu, oracle, m, n):
    rs = np.random.RandomState(m * 7 + n)
    y = rs.standard_normal(n + m); p, q = weights(m, n, 2, False)
    assert np.array_equal(gpu.Aty(y, p, q), oracle.Aty(y, p, q))


def test_ax_aty_device_tensors_and_adjointness(gpu):
    import torch
    m, n = 1536, 1100
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(m * n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(n + m, dtype=torch.float64, device="cuda", generator=g)
    p = torch.rand(m, dtype=torch.float64, device="cuda", generator=g) + 0.5
    q = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 0.5
    ax = gpu.Ax(x, p, q); aty = gpu.Aty(y, p, q)
    assert ax.is_cuda and aty.is_cuda
    lhs = float(ax @ y); rhs = float(x @ aty)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), 1.0)


def test_ax_equals_explicit_matrix(gpu, oracle):
    m, n = 23, 17
    rs = np.random.RandomState(0)
    p, q = weights(m, n, 3, False)
    A = oracle.explicit_A(p, q)                      # Class1/APD_SsN_Class1.m:47
    x = rs.standard_normal(m * n); y = rs.standard_normal(m + n)
    assert close(gpu.Ax(x, p, q), A @ x)
    assert close(gpu.Aty(y, p, q), A.T @ y)


@pytest.mark.parametrize("m,n", [(7, 5), (128, 64), (500, 500), (1027, 517), (2048, 300)])
@pytest.mark.parametrize("gmode", ["inf", "scalar", "vector"])
def test_prox_residual_matches_oracle(gpu, oracle, m, n, gmode):
    rs = np.random.RandomState(m + 3 * n)
    p, q = weights(m, n, 4, False)
    w = rs.standard_normal(m * n); lam = 0.5 * rs.standard_normal(m + n); tk = 0.37
    gama = {"inf": np.inf, "scalar": 0.8, "vector": rs.random_sample(m * n) + 0.1}[gmode]
    z = 1 / tk * (w - oracle.Aty(lam, p, q))         # Class1/APD_SsN_Class1.m:139
    s = (z >= 0) & (z <= gama)                       # :140
    px = np.minimum(np.maximum(0.0, z), gama)        # :32
    out = gpu.prox_residual(w, lam, p, q, tk, gama, want=("Axprox", "prox", "z", "s"))
    assert np.array_equal(out["z"], z)
    assert np.array_equal(out["s"].astype(bool), s)
    assert np.array_equal(out["prox"], px)
    assert out["count"] == int(s.sum())
    assert close(out["Axprox"], oracle.Ax(px, p, q))
    # the line-search term of APD_SsN_Class1.m:183-187: prob < 3 (gama = Inf) and prob = 3 (capacities)
    term = float(px @ px) if gmode == "inf" else float(z @ z) - float((z - px) @ (z - px))
    assert abs(out["norm2"] - term) <= 1e-12 * float(z @ z) + 1e-300
    lite = gpu.prox_residual(w, lam, p, q, tk, gama, want=())      # line-search form: norm only
    assert abs(lite["norm2"] - out["norm2"]) <= 1e-14 * abs(out["norm2"])


@pytest.mark.parametrize("m,n", [(7, 5), (128, 64), (500, 500), (1027, 517), (2048, 300)])
@pytest.mark.parametrize("unit_phi", [True, False])
def test_prox_residual_pot_matches_oracle(gpu, oracle, m, n, unit_phi):
    """Fused residual of partial OT (Class2/APD_SsN_Class2.m:124-130,137-150): z, the active flags s and t and prox(z) bit
    for bit against the reference expression, H*prox(z) to 1e-10."""
    rs = np.random.RandomState(5 * m + n)
    p, q = weights(m, n, 6, False)
    N = m + n; mn = m * n
    w = rs.standard_normal(mn + N); lam = 0.5 * rs.standard_normal(N + 1); tk = 0.41
    phi = np.ones(mn) if unit_phi else rs.random_sample(mn) + 0.5
    Htlk = np.concatenate([oracle.Aty(lam[:N], p, q) + lam[N] * phi, lam[:N]])      # :124
    z = 1 / tk * (w - Htlk)                                                        # :127
    pz = np.maximum(z, 0.0)
    Hp = np.concatenate([oracle.Ax(pz[:mn], p, q) + pz[mn:], [phi @ pz[:mn]]])      # :128-129
    out = gpu.prox_residual_pot(w, lam, p, q, tk, phi, want=("Hprox", "prox", "s", "t"))
    assert np.array_equal(out["s"].astype(bool), z[:mn] >= 0)
    assert np.array_equal(out["t"] > 0.5, z[mn:] >= 0)
    assert np.array_equal(out["prox"], pz)
    assert out["count"] == int((z[:mn] >= 0).sum())
    assert close(out["Hprox"], Hp)
    assert abs(out["norm2"] - float(pz @ pz)) <= 1e-12 * float(pz @ pz)
    lite = gpu.prox_residual_pot(w, lam, p, q, tk, phi, want=())                   # line-search form: the norm only
    assert abs(lite["norm2"] - out["norm2"]) <= 1e-14 * abs(out["norm2"])


@pytest.mark.parametrize("m,n,density", [(5, 4, 0.5), (64, 48, 0.1), (300, 500, 0.02), (512, 512, 0.004),
                                           (1000, 37, 0.3), (130, 2100, 0.01)])
@pytest.mark.parametrize("unit", [True, False])
def test_asat_pattern_and_values_exact(gpu, oracle, m, n, density, unit):
    s, p, q = random_active_problem(m, n, density, seed=m + n, weights=not unit)
    H_ref = oracle.ASAt(s, p, q)
    H = gpu.ASAt(s, p, q).to_scipy().tocsc(); H.sort_indices()
    assert H.shape == (m + n, m + n)
    assert np.array_equal(H.indptr, H_ref.indptr)
    assert np.array_equal(H.indices, H_ref.indices)
    assert np.array_equal(H.data, H_ref.data)
    # ASAt == A*diag(s)*A'  (ASAt.m:3-12)
    A = oracle.explicit_A(p, q)
    import scipy.sparse as sp
    E = (A @ sp.diags(s.astype(float)) @ A.T).tocsc()
    assert abs(E - H).max() <= 1e-10 * max(abs(E).max(), 1.0)


def test_asat_edge_cases(gpu, oracle):
    m, n = 40, 24
    p, q = np.ones(m), np.ones(n)
    H = gpu.ASAt(np.zeros(m * n, dtype=bool), p, q)
    assert H.nnz == 0 and H.shape == (m + n, m + n)
    s = np.ones(m * n, dtype=bool)
    Hs = gpu.ASAt(s, p, q).to_scipy().tocsc(); Hs.sort_indices()
    R = oracle.ASAt(s, p, q)
    assert np.array_equal(Hs.indices, R.indices) and np.array_equal(Hs.data, R.data)
    s = np.zeros((m, n), dtype=bool); s[3, :] = True; s[:, 5] = True      # isolated rows/cols elsewhere
    Hs = gpu.ASAt(s.reshape(-1, order="F"), p, q).to_scipy().tocsc(); Hs.sort_indices()
    R = oracle.ASAt(s.reshape(-1, order="F"), p, q)
    assert np.array_equal(Hs.indptr, R.indptr) and np.array_equal(Hs.indices, R.indices)


def test_asatz(gpu, oracle):
    m = n = 48
    s, p, q = random_active_problem(m, n, 0.1, seed=9, weights=True)
    z = np.random.RandomState(1).standard_normal(m + n)
    assert close(gpu.ASAtz(z, s, p, q), oracle.ASAtz(z, s, p, q))
    # with p == q the reference's Q*p typo is harmless: ASAtz == ASAt*z   (ASAtz.m:10-13)
    assert close(gpu.ASAtz(z, s, p, p), oracle.ASAt(s, p, p) @ z)
    with pytest.raises(gpu.SsnError) as ei:
        gpu.ASAtz(np.zeros(7), np.zeros(12, dtype=bool), np.ones(3), np.ones(4))
    assert ei.value.status == "SSN_E_ASATZ_DIM"


def test_invaat_invhht(gpu, oracle):
    import scipy.sparse as sp
    m, n = 31, 22
    rs = np.random.RandomState(3)
    p, q = weights(m, n, 5, False)
    x = rs.standard_normal(m + n)
    for args in [(), (0.7,), (0.7, 1.9)]:
        assert close(gpu.invAAt(x, p, q, *args), oracle.invAAt(x, p, q, *args))
    A = oracle.explicit_A(p, q)
    y = gpu.invAAt(x, p, q, 0.7, 1.9)
    M = sp.diags(np.concatenate([0.7 * np.ones(n), 1.9 * np.ones(m)])) + A @ A.T          # invAAt.m:2
    assert close(M @ y, x, 1e-9)
    phi = rs.random_sample(m * n); v = rs.standard_normal(m + n + 1)
    assert close(gpu.invHHt(v, p, q, 0.6, phi), oracle.invHHt(v, p, q, 0.6, phi))


def test_matlab_random_stream(gpu, oracle):
    first = np.load(os.path.join(GOLDEN, "matlab_rand_first.npz"))["first"]
    # MATLAB's documented first draws after start-up
    assert np.allclose(first[:5], [0.8147236863931789, 0.9057919370756192, 0.12698681629350606,
                                   0.9133758561390194, 0.6323592462254095], rtol=0, atol=1e-15)
    gpu.rng_reset()
    a = gpu.rand(313).cpu().numpy(); b = gpu.rand(687).cpu().numpy()          # state carries across calls
    assert np.array_equal(np.concatenate([a, b]), first)
    assert gpu.rng_drawn() == 1000
    gpu.rng_reset(); oracle.rng_reset()
    assert np.array_equal(gpu.rand(5000).cpu().numpy(), oracle.rand(5000))


def test_full_size_properties_64x64_grid(gpu):
    """Config 2 size (m=n=4096, 16.8M-entry plan): size-independent identities."""
    import torch
    m = n = 4096
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(m * n, dtype=torch.float64, device="cuda", generator=g)
    ones_m = torch.ones(m, dtype=torch.float64, device="cuda"); ones_n = torch.ones(n, dtype=torch.float64, device="cuda")
    y = gpu.Ax(x, ones_m, ones_n)
    X = x.view(n, m)                                  # column-major m x n == row-major n x m
    assert torch.allclose(y[:n], X.sum(dim=1), rtol=1e-11, atol=0)
    assert torch.allclose(y[n:], X.sum(dim=0), rtol=1e-11, atol=0)
    assert abs(float(y[:n].sum()) - float(y[n:].sum())) <= 1e-10 * float(y.sum())     # checksum of checksums
    lam = torch.randn(m + n, dtype=torch.float64, device="cuda", generator=g)
    z = gpu.Aty(lam, ones_m, ones_n)
    assert torch.equal(z.view(n, m), lam[:n, None] + lam[None, n:])                   # rank-2 structure
    lhs = float(gpu.Ax(x, ones_m, ones_n) @ lam); rhs = float(x @ z)
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
    # fused residual == composition of the plain operators
    out = gpu.prox_residual(x - 0.5, lam, ones_m, ones_n, 0.9, float("inf"), want=("Axprox", "s", "prox"))
    zz = (1 / 0.9) * ((x - 0.5) - z)
    assert torch.equal(out["s"].bool(), zz >= 0)
    assert torch.equal(out["prox"], torch.clamp(zz, min=0.0))
    assert torch.allclose(out["Axprox"], gpu.Ax(out["prox"], ones_m, ones_n), rtol=1e-11, atol=0)
    H = gpu.ASAt(out["s"], ones_m, ones_n).to_scipy()
    assert H.nnz == 2 * out["count"] + np.count_nonzero(H.diagonal())
    assert abs(H - H.T).max() == 0
    rowsum = np.asarray(H.sum(axis=1)).reshape(-1)
    assert np.array_equal(rowsum, 2 * H.diagonal())   # unit weights: diag = degree, off-diag ones


@pytest.mark.parametrize("m,n", [(7, 5), (250, 130), (1027, 517), (2048, 300)])
@pytest.mark.parametrize("gama", [np.inf, 0.4])
def test_prox_trials_equal_single_trial_kernel(gpu, m, n, gama):
    """The batched line-search kernel (up to 8 trial vectors per read of w) against the fused
    single-trial residual kernel: same per-entry arithmetic, same reduction order => same bits."""
    rs = np.random.RandomState(m + 3 * n)
    w = rs.standard_normal(m * n); p, q = weights(m, n, 3, False)
    for nt in (1, 2, 3, 4, 5, 7, 8):
        lamT = 0.3 * rs.standard_normal((nt, n + m))
        got = gpu.prox_trials(w, lamT, p, q, 0.8, gama).cpu().numpy()
        ref = np.array([gpu.prox_residual(w, lamT[t], p, q, 0.8, gama, want=())["norm2"] for t in range(nt)])
        assert np.array_equal(got, ref), (nt, got, ref)


@pytest.mark.parametrize("m,n", [(7, 5), (250, 130), (1027, 517), (2048, 300)])
@pytest.mark.parametrize("unit", [True, False])
@pytest.mark.parametrize("shift", [0.0, 3.0])
def test_screened_trials_equal_dense_trials(gpu, m, n, unit, shift):
    """The screened path (up to 128 backtracking steps of one direction per read of w: entries whose
    first and last trial residuals are safely negative are dropped, the others are listed and
    evaluated with the literal per-step expression) against the dense batched kernel on the same trial
    vectors.  Only entries that add exactly 0 are dropped, so the two differ by the summation order
    alone: <= 1e-13 relative.  shift = 3 makes the trial plans sparse (few entries with z > 0), the
    regime the screen is for; shift = 0 keeps about half of the entries active."""
    rs = np.random.RandomState(5 * m + n)
    w = rs.standard_normal(m * n) - shift
    p, q = (np.ones(m), np.ones(n)) if unit else weights(m, n, 3, False)
    lam = 0.3 * rs.standard_normal(n + m); zeta = 0.5 * rs.standard_normal(n + m); wlk = rs.standard_normal(n + m)
    total = float(m * n)
    for ll0, nt in ((0, 1), (0, 8), (3, 5), (1, 16), (40, 13), (0, 32), (7, 27), (2, 70), (0, 128)):
        got = gpu.prox_trials_lin(w, lam, zeta, p, q, 0.8, 0.9, ll0, nt).cpu().numpy()
        again = gpu.prox_trials_lin(w, lam, zeta, p, q, 0.8, 0.9, ll0, nt).cpu().numpy()
        assert np.array_equal(got, again)                      # deterministic
        ref = []
        for t0 in range(0, nt, 8):
            k = min(8, nt - t0)
            lamT, _ = gpu.trial_vectors(lam, zeta, wlk, 0.9, ll0 + t0, k)
            ref.append(gpu.prox_trials(w, lamT, p, q, 0.8, np.inf).cpu().numpy())
        ref = np.concatenate(ref)
        assert np.all(np.abs(got[:nt] - ref) <= 1e-13 * np.abs(ref)), (ll0, nt, got[:nt], ref)
        assert 0 <= got[nt] <= total
    if shift > 0 and m * n > 10000:
        assert got[nt] < 0.1 * total                  # the screen does drop most entries when the plan is sparse


def test_screened_trials_count_exact_active_entries(gpu):
    """With a single step per batch the first and the last trial coincide, so the candidates are the
    entries with z > -tol: every active entry (z >= 0) is among them and almost nothing else."""
    m, n = 700, 420
    rs = np.random.RandomState(2)
    w = rs.standard_normal(m * n) - 2.0; p, q = np.ones(m), np.ones(n)
    lam = 0.3 * rs.standard_normal(n + m); zeta = 0.5 * rs.standard_normal(n + m)
    out = gpu.prox_trials_lin(w, lam, zeta, p, q, 0.8, 0.9, 4, 1).cpu().numpy()
    ev = gpu.prox_residual(w, lam + 0.9 ** 4 * zeta, p, q, 0.8, np.inf, want=())
    assert ev["count"] <= out[1] <= ev["count"] + 4
    assert abs(out[0] - ev["norm2"]) <= 1e-13 * ev["norm2"]


def test_adaptive_linesearch_equals_fixed_batches(gpu):
    """ssn_linesearch with batch = 0 (screened path, 32/64/128 steps per pass) accepts the same step as
    batch = 8 (dense kernel): the objective values agree to rounding, so the Armijo decisions agree
    wherever they are not decided by the last bits."""
    import torch
    m, n = 640, 520
    rs = np.random.RandomState(11)
    p, q = np.ones(m), np.ones(n)
    w = rs.standard_normal(m * n) - 2.5; lam = 0.2 * rs.standard_normal(n + m); wlk = rs.standard_normal(n + m)
    tk, bk1, nu, delta = 0.7, 0.3, 0.2, 0.9
    ev = gpu.prox_residual(w, lam, p, q, tk, np.inf, want=("Axprox",))
    axp = ev["Axprox"]
    grad = bk1 * lam - wlk - (axp.cpu().numpy() if hasattr(axp, "cpu") else np.asarray(axp))
    cF_old = bk1 / 2 * (lam @ lam) - wlk @ lam + 0.5 * tk * ev["norm2"]
    for scale, ll_max in ((0.5, 500), (300.0, 500), (5000.0, 500), (-30.0, 45), (-30.0, 230)):
        zeta = -scale * grad
        ress = abs(float(grad @ zeta))
        a = gpu.linesearch(w, lam, zeta, wlk, p, q, tk, bk1, cF_old, ress, np.inf, nu, delta, ll_max, batch=8)
        b = gpu.linesearch(w, lam, zeta, wlk, p, q, tk, bk1, cF_old, ress, np.inf, nu, delta, ll_max, batch=0)
        assert a[1] == b[1], (scale, a[1:], b[1:])
        assert abs(a[2] - b[2]) <= 1e-13 * abs(a[2]) and abs(a[3] - b[3]) <= 1e-13 * max(1.0, abs(a[3]))
        assert torch.equal(a[0], b[0])
        assert b[4] <= a[4]


def test_linesearch_matches_trial_by_trial_loop(gpu, oracle):
    """ssn_linesearch (ll = 0 alone, then 8 trials per pass) against the reference's trial-by-trial Armijo loop
    (Class1/APD_SsN_Class1.m:182-211) evaluated with the oracle."""
    m, n = 300, 260
    rs = np.random.RandomState(9)
    p, q = np.ones(m), np.ones(n)
    w = rs.standard_normal(m * n) - 0.5; lam = 0.2 * rs.standard_normal(n + m); wlk = rs.standard_normal(n + m)
    tk, bk1, nu, delta = 0.7, 0.3, 0.2, 0.9
    prox = lambda z: np.maximum(z, 0.0)
    cF = lambda l: bk1 / 2 * (l @ l) - wlk @ l + 0.5 * tk * np.sum(prox((w - oracle.Aty(l, p, q)) / tk) ** 2)
    z0 = (w - oracle.Aty(lam, p, q)) / tk
    grad = bk1 * lam - wlk - oracle.Ax(prox(z0), p, q)               # gradient of cF at lam
    for scale, ll_max in ((0.5, 500), (50.0, 500), (2000.0, 500), (-30.0, 7)):
        zeta = -scale * grad                                         # descent direction (ascent for scale < 0)
        cF_old = cF(lam); ress = abs(float(grad @ zeta))
        ll = 0
        while True:
            lk_new = lam + delta ** ll * zeta
            if not (cF(lk_new) > cF_old - nu * delta ** ll * ress) or ll == ll_max:
                break
            ll += 1
        out, ll_dev, n2, cF_new, passes = gpu.linesearch(w, lam, zeta, wlk, p, q, tk, bk1, cF_old, ress, np.inf, nu, delta, ll_max, batch=8)
        assert ll_dev == ll, (scale, ll_dev, ll)
        assert passes == 1 + (ll + 7) // 8
        assert np.array_equal(out.cpu().numpy(), lk_new)
        assert abs(cF_new - cF(lk_new)) <= 1e-10 * max(1.0, abs(cF_new))


@pytest.mark.parametrize("gama", [0.8, "vector"])
def test_linesearch_with_capacities_uses_the_prob3_merit(gpu, oracle, gama):
    """Finite gama (prob = 3): the Armijo function is f0 + tk/2*(||z||^2 - ||z - prox(z)||^2)
    (Class1/APD_SsN_Class1.m:185-186,194-197,203-207), not tk/2*||prox(z)||^2."""
    m, n = 130, 90
    rs = np.random.RandomState(11)
    p, q = np.ones(m), np.ones(n)
    w = rs.standard_normal(m * n) + 0.3; lam = 0.2 * rs.standard_normal(n + m); wlk = rs.standard_normal(n + m)
    if isinstance(gama, str):
        gama = rs.random_sample(m * n) + 0.2
    tk, bk1, nu, delta, ll_max = 0.7, 0.3, 0.2, 0.9, 500
    zof = lambda l: (w - oracle.Aty(l, p, q)) / tk
    prox = lambda z: np.minimum(np.maximum(z, 0.0), gama)
    cF = lambda l: bk1 / 2 * (l @ l) - wlk @ l + 0.5 * tk * (np.sum(zof(l) ** 2) - np.sum((zof(l) - prox(zof(l))) ** 2))
    grad = bk1 * lam - wlk - oracle.Ax(prox(zof(lam)), p, q)          # gradient of cF at lam
    for scale in (0.5, 40.0, 900.0):
        zeta = -scale * grad
        cF_old = cF(lam); ress = abs(float(grad @ zeta))
        ll = 0
        while True:
            lk_new = lam + delta ** ll * zeta
            if not (cF(lk_new) > cF_old - nu * delta ** ll * ress) or ll == ll_max:
                break
            ll += 1
        ev = gpu.prox_residual(w, lam, p, q, tk, gama, want=())
        assert abs(bk1 / 2 * (lam @ lam) - wlk @ lam + 0.5 * tk * ev["norm2"] - cF_old) <= 1e-11 * abs(cF_old)
        out, ll_dev, n2, cF_new, passes = gpu.linesearch(w, lam, zeta, wlk, p, q, tk, bk1, cF_old, ress, gama, nu, delta, ll_max)
        assert ll_dev == ll, (scale, ll_dev, ll)
        assert np.array_equal(out.cpu().numpy(), lk_new)
        assert abs(cF_new - cF(lk_new)) <= 1e-10 * max(1.0, abs(cF_new))


@pytest.mark.parametrize("m,n", [(7, 5), (250, 130), (1027, 517), (2048, 300)])
@pytest.mark.parametrize("gama", [np.inf, 0.3])
def test_fused_apd_outer_updates(gpu, oracle, m, n, gama):
    """ssn_apd_begin / ssn_apd_end against the reference expressions of
    Class1/APD_SsN_Class1.m:125-126 and :239-254 evaluated with the oracle operators."""
    rs = np.random.RandomState(3 * m + n)
    p, q = weights(m, n, 4, False)
    c = rs.random_sample(m * n); xk = np.maximum(rs.standard_normal(m * n), 0); vk = rs.standard_normal(m * n)
    lam = 0.4 * rs.standard_normal(n + m)
    ak, bk, tk = 1.7, 0.6, 0.45
    prox = lambda x: np.minimum(np.maximum(0.0, x), gama)
    wk_ref = -c + bk * (xk + ak * vk) / ak ** 2
    wk, axk = gpu.apd_begin(c, xk, vk, p, q, ak, bk)
    assert np.array_equal(wk.cpu().numpy(), wk_ref)
    assert close(axk.cpu().numpy(), oracle.Ax(xk, p, q))
    zk = 1 / tk * (wk_ref - oracle.Aty(lam, p, q))
    x1_ref = prox(zk); v1_ref = x1_ref + (x1_ref - xk) / ak
    kx_ref = np.linalg.norm(x1_ref - prox(x1_ref - c - oracle.Aty(lam, p, q)))
    x1, v1, ax1, cx, kx2 = gpu.apd_end(c, wk, xk, lam, p, q, tk, ak, gama)
    assert np.array_equal(x1.cpu().numpy(), x1_ref) and np.array_equal(v1.cpu().numpy(), v1_ref)
    assert close(ax1.cpu().numpy(), oracle.Ax(x1_ref, p, q))
    assert abs(cx - c @ x1_ref) <= 1e-10 * max(abs(c @ x1_ref), 1.0)
    assert abs(np.sqrt(kx2) - kx_ref) <= 1e-10 * max(kx_ref, 1.0)
