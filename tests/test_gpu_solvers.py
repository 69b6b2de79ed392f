"""GPU parity tests, inner linear-solve dispatch: components, rescaled system, Hybrid_AMG,
aug_PCG, AMG4POT, PCG4POT -- against the CPU oracle and against the committed golden systems
(real SsN systems recorded from the reference's bundled input and from a grid problem)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_systems, load_system, random_active_problem
from test_gpu_amg import AMG_OPTS, assert_same_matrix, csc_sorted

pytestmark = pytest.mark.gpu


def prob(oracle, m, n, density, seed, weights=False, bk1=0.05, tk=0.8, tdiag=None, connect=True):
    s, p, q = random_active_problem(m, n, density, seed, weights)
    if not connect:
        S = s.reshape((m, n), order="F").copy()
        S[: m // 2, n // 2:] = False; S[m // 2:, : n // 2] = False          # two big blocks
        S[0, :] = False; S[:, 0] = False; S[0, 0] = True                     # a 2-node component
        S[1, :] = False                                                      # an isolated row node
        S[2:6, :] = False; S[:, 2:5] = False; S[2:6, 2:5] = True             # a small dense component
        s = S.reshape(-1, order="F")
    H0 = oracle.ASAt(s, p, q)
    t = np.zeros(m + n) if tdiag is None else tdiag
    return {"bk1": bk1, "tk": tk, "p": p, "q": q, "T": sp.diags(t), "H0": H0,
            "z": np.random.RandomState(seed + 1).standard_normal(m + n), "s": s}


def test_components_ordering(gpu, oracle):
    for seed, n, d in [(0, 60, 0.02), (1, 500, 0.002), (2, 300, 0.01)]:
        rs = np.random.RandomState(seed)
        A = sp.random(n, n, density=d, random_state=rs, format="csr"); A = (A + A.T + sp.identity(n)).tocsr()
        b_ref, s_ref, p_ref, r_ref = oracle.components(A)
        b, s, p, r = gpu.components(A)
        assert np.array_equal(b, b_ref) and np.array_equal(s, s_ref)
        assert np.array_equal(p, p_ref) and np.array_equal(r, r_ref)
    # a long path: label propagation must still converge (pointer jumping)
    n = 3000
    P = sp.diags([np.ones(n - 1), np.ones(n - 1), 2 * np.ones(n)], [-1, 1, 0], format="csr")
    b, s, p, r = gpu.components(P)
    assert s.tolist() == [n] and np.array_equal(p, np.arange(n))
    with pytest.raises(gpu.SsnError) as ei:
        gpu.components(sp.random(4, 5, density=0.5, format="csr"))
    assert ei.value.status == "SSN_E_NOT_SQUARE"


@pytest.mark.parametrize("weights", [False, True])
def test_rescaled_system_bit_exact(gpu, oracle, weights):
    from oracle.solvers import rescaled_system
    m, n = 120, 90
    t = np.random.RandomState(0).random_sample(m + n) * (np.arange(m + n) % 3 == 0)
    pd = prob(oracle, m, n, 0.04, 11, weights, tdiag=t)
    qp, A0, Qd, Kd, Ae_ref, f_ref = rescaled_system(pd)
    Ae, f = gpu.rescaled_system(pd)
    assert_same_matrix(Ae, Ae_ref, what="Ae")
    assert np.array_equal(f, f_ref)


@pytest.mark.parametrize("m,n,density,weights", [(60, 50, 0.06, False), (400, 300, 0.01, True), (1200, 1000, 0.003, False)])
def test_hybrid_amg_connected(gpu, oracle, m, n, density, weights):
    pd = prob(oracle, m, n, density, 21 + m, weights)
    oracle.rng_reset(); gpu.rng_reset()
    z_ref, it_ref, res_ref, info_ref = oracle.Hybrid_AMG(pd, AMG_OPTS)
    zeta, it, res, info = gpu.Hybrid_AMG(pd, AMG_OPTS)
    assert list(info) == list(info_ref)
    assert it == it_ref
    assert gpu.rng_drawn() == oracle.GLOBAL_STREAM.drawn
    Jk = pd["bk1"] * sp.identity(m + n) + pd["H0"] / pd["tk"]              # Class1/APD_SsN_Class1.m:143
    assert np.linalg.norm(Jk @ zeta - pd["z"]) <= 1e-9 * np.linalg.norm(pd["z"])
    assert np.linalg.norm(zeta - z_ref) <= 1e-7 * np.linalg.norm(z_ref)


def test_hybrid_amg_disconnected_large_and_small(gpu, oracle):
    m, n = 420, 380
    pd = prob(oracle, m, n, 0.02, 5, connect=False)
    oracle.rng_reset(); gpu.rng_reset()
    z_ref, it_ref, res_ref, info_ref = oracle.Hybrid_AMG(pd, AMG_OPTS)
    zeta, it, res, info = gpu.Hybrid_AMG(pd, AMG_OPTS)
    assert info_ref[0] > 3
    assert list(info) == list(info_ref) and it == it_ref
    assert gpu.rng_drawn() == oracle.GLOBAL_STREAM.drawn
    Jk = pd["bk1"] * sp.identity(m + n) + pd["H0"] / pd["tk"]
    assert np.linalg.norm(Jk @ zeta - pd["z"]) <= 1e-9 * np.linalg.norm(pd["z"])
    assert np.linalg.norm(zeta - z_ref) <= 1e-7 * np.linalg.norm(z_ref)


def test_hybrid_amg_all_isolated(gpu, oracle):
    # E = 0: every node is its own component, pure diagonal solve (first SsN step of outer it 2)
    m, n = 50, 40
    pd = {"bk1": 0.3, "tk": 0.7, "p": np.ones(m), "q": np.ones(n), "T": sp.diags(np.zeros(m + n)),
          "H0": oracle.ASAt(np.zeros(m * n, dtype=bool), np.ones(m), np.ones(n)),
          "z": np.random.RandomState(0).standard_normal(m + n)}
    z_ref, it_ref, res_ref, info_ref = oracle.Hybrid_AMG(pd, AMG_OPTS)
    zeta, it, res, info = gpu.Hybrid_AMG(pd, AMG_OPTS)
    assert list(info) == list(info_ref) == [m + n, 0] and it == 0
    assert np.allclose(zeta, z_ref, rtol=1e-14)


def test_hybrid_amg_spd_branch_with_T(gpu, oracle):
    # Class2: T = diag(t) nonzero -> isnsp = 0 (Hybrid_AMG.m:32-35)
    m, n = 200, 180
    t = (np.random.RandomState(3).random_sample(m + n) > 0.5).astype(float)
    pd = prob(oracle, m, n, 0.03, 9, tdiag=t)
    oracle.rng_reset(); gpu.rng_reset()
    z_ref, it_ref, _, info_ref = oracle.Hybrid_AMG(pd, AMG_OPTS)
    zeta, it, _, info = gpu.Hybrid_AMG(pd, AMG_OPTS)
    assert list(info) == list(info_ref) and it == it_ref
    assert np.linalg.norm(zeta - z_ref) <= 1e-7 * np.linalg.norm(z_ref)


def test_pq_zero_error(gpu, oracle):
    m, n = 20, 15
    pd = prob(oracle, m, n, 0.1, 2)
    pd["q"] = pd["q"].copy(); pd["q"][3] = 0.0
    with pytest.raises(gpu.SsnError) as ei:
        gpu.Hybrid_AMG(pd, AMG_OPTS)
    assert ei.value.status == "SSN_E_PQ_ZERO"                                # Hybrid_AMG.m:18-19


@pytest.mark.parametrize("path", golden_systems("bundled500") + golden_systems("grid12"))
def test_golden_ssn_systems(gpu, path):
    """Real SsN systems (recorded by tests/golden/make_golden.py): ASAt pattern, hierarchy sizes,
    C/F vectors, cycle counts and solution against the committed oracle outputs."""
    d = load_system(path)
    m, n = d["m"], d["n"]
    p, q = np.ones(m), np.ones(n)
    H = gpu.ASAt(d["s"], p, q)
    Hs = csc_sorted(H.to_scipy())
    assert np.array_equal(Hs.indptr, d["H_indptr"]) and np.array_equal(Hs.indices, d["H_indices"])
    assert np.array_equal(Hs.data, d["H_data"])
    pd = {"bk1": float(d["bk1"]), "tk": float(d["tk"]), "p": p, "q": q, "T": None, "H0": H, "z": d["z"]}
    gpu.rng_reset()
    zeta, it, res, info = gpu.Hybrid_AMG(pd, AMG_OPTS)
    assert list(info) == list(d["info"])
    assert it == int(d["it"])
    assert gpu.rng_drawn() == int(d["drawn"])
    assert np.linalg.norm(zeta - d["zeta"]) <= 1e-6 * np.linalg.norm(d["zeta"])
    Jk = pd["bk1"] * sp.identity(m + n) + Hs / pd["tk"]
    assert np.linalg.norm(Jk @ zeta - d["z"]) <= 1e-9 * np.linalg.norm(d["z"])
    if "level_sizes" in d:
        Ae, f = gpu.rescaled_system(pd)
        gpu.rng_reset()
        gpu.rand(m + n)                                  # the initial guess is drawn first (Hybrid_AMG.m:40)
        levels = gpu.amg_setup(Ae, dict(AMG_OPTS, fnode=n, isnsp=1))
        assert [a.shape[0] for a, _ in levels] == d["level_sizes"].tolist()
        assert [a.nnz for a, _ in levels] == d["level_nnz"].tolist()
        for k in range(2, len(levels)):
            isC = np.unpackbits(d[f"isC_{k}"])[: levels[k - 1][0].shape[0]].astype(bool)
            Pk = levels[k][1].to_scipy().tocsr()
            unit_rows = (np.diff(Pk.indptr) == 1) & (Pk.data[np.minimum(Pk.indptr[:-1], Pk.nnz - 1)] == 1.0)
            assert np.all(unit_rows[isC]), f"C nodes of level {k} differ from the golden split"
        gpu.amg_clear()


def test_aug_pcg(gpu, oracle):
    m, n = 150, 130
    for connect in (True, False):
        pd = prob(oracle, m, n, 0.03, 31, bk1=0.2, connect=connect)
        o = {"retol": 1e-11, "maxit": 10000, "precd": 2, "guess": None}
        z_ref, it_ref, res_ref, info_ref = oracle.aug_PCG(pd, o)
        zeta, it, res, info = gpu.aug_PCG(pd, o)
        assert list(info) == list(info_ref)
        assert abs(it - it_ref) <= max(3, it_ref // 20)
        assert np.linalg.norm(zeta - z_ref) <= 1e-7 * np.linalg.norm(z_ref)


def test_pot_bordered_solves(gpu, oracle):
    m, n = 140, 120
    t = (np.random.RandomState(4).random_sample(m + n) > 0.4).astype(float)
    pd = prob(oracle, m, n, 0.03, 41, bk1=0.2, tdiag=t)
    pd["phi"] = np.random.RandomState(5).random_sample(m * n) + 0.5
    pd["z"] = np.random.RandomState(6).standard_normal(m + n + 1)
    oracle.rng_reset(); gpu.rng_reset()
    z_ref, it_ref, res_ref, info_ref = oracle.AMG4POT(pd, dict(AMG_OPTS, maxit=40, smoth=10), "amg")
    zeta, it, res, info = gpu.AMG4POT(pd, dict(AMG_OPTS, maxit=40, smoth=10), "amg")
    assert list(info) == list(info_ref) and it == it_ref
    assert np.linalg.norm(zeta - z_ref) <= 1e-7 * np.linalg.norm(z_ref)
    # the bordered system itself (Class2/AMG4POT.m:4-10)
    s = pd["s"].astype(float); A = oracle.explicit_A(pd["p"], pd["q"])
    ss = A @ (s * pd["phi"])
    cH = sp.bmat([[pd["T"] + pd["H0"], ss[:, None]], [ss[None, :], [[pd["phi"] @ (s * pd["phi"])]]]])
    He = pd["bk1"] * sp.identity(m + n + 1) + cH / pd["tk"]
    assert np.linalg.norm(He @ zeta - pd["z"]) <= 1e-8 * np.linalg.norm(pd["z"])
    o = {"retol": 1e-11, "maxit": 10000, "precd": 2, "guess": None}
    z2_ref, *_ = oracle.PCG4POT(pd, o)
    z2, *_ = gpu.PCG4POT(pd, o)
    assert np.linalg.norm(z2 - z2_ref) <= 1e-7 * np.linalg.norm(z2_ref)
    # str = 'twogrid' (Class2/AMG4POT.m:48-51; inner_solver = 5 of the Class 2 script): the two solves through Hybrid_twogrid
    oracle.rng_reset(); gpu.rng_reset()
    z3_ref, it3_ref, _, info3_ref = oracle.AMG4POT(pd, dict(AMG_OPTS, maxit=40, smoth=10), "twogrid")
    z3, it3, _, info3 = gpu.AMG4POT(pd, dict(AMG_OPTS, maxit=40, smoth=10), "twogrid")
    assert list(info3) == list(info3_ref) and it3 == it3_ref
    assert np.linalg.norm(z3 - z3_ref) <= 1e-7 * np.linalg.norm(z3_ref)
    assert np.linalg.norm(He @ z3 - pd["z"]) <= 1e-8 * np.linalg.norm(pd["z"])


@pytest.mark.parametrize("tag", ["k12_s2", "k30_s1", "k40_s2", "k80_s2"])
def test_full_size_128x128_grid_systems(gpu, oracle, tag):
    """SsN systems of the headline configuration (128x128 grid, N = 32768) recorded from the device
    solve (tools/save_states.py): ASAt from the recorded active set, then Hybrid_AMG against the
    oracle -- identical component count, cycle count, random draws; solution <= 1e-6."""
    import os
    import torch
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "ssn_states_g128.npz"))
    g = int(d["g"]); m = n = g * g
    lin = d[tag + "_lin"]
    s = torch.zeros(m * n, dtype=torch.uint8, device="cuda"); s[torch.from_numpy(lin).cuda()] = 1
    p = np.ones(m); q = np.ones(n)
    H = gpu.ASAt(s, torch.ones(m, dtype=torch.float64, device="cuda"), torch.ones(n, dtype=torch.float64, device="cuda"))
    del s
    ii, jj = lin % m, lin // m                       # the same matrix for the oracle, from the coordinates
    rows = np.concatenate([np.arange(n), jj, n + ii, n + np.arange(m)]); cols = np.concatenate([np.arange(n), n + ii, jj, n + np.arange(m)])
    vals = np.concatenate([np.bincount(jj, minlength=n).astype(float), np.ones(lin.size), np.ones(lin.size), np.bincount(ii, minlength=m).astype(float)])
    H_ref = sp.csc_matrix((vals, (rows, cols)), shape=(m + n, m + n)); H_ref.eliminate_zeros(); H_ref.sort_indices()
    Hs = csc_sorted(H.to_scipy())
    assert np.array_equal(Hs.indptr, H_ref.indptr) and np.array_equal(Hs.indices, H_ref.indices) and np.array_equal(Hs.data, H_ref.data)
    pd = {"bk1": float(d[tag + "_bk1"]), "tk": float(d[tag + "_tk"]), "p": p, "q": q, "T": sp.diags(np.zeros(m + n)), "H0": H_ref, "z": d[tag + "_z"]}
    oracle.rng_reset(); gpu.rng_reset()
    z_ref, it_ref, res_ref, info_ref = oracle.Hybrid_AMG(pd, AMG_OPTS)
    zeta, it, res, info = gpu.Hybrid_AMG(dict(pd, H0=H), AMG_OPTS)
    assert list(info) == list(info_ref) and it == it_ref
    assert gpu.rng_drawn() == oracle.GLOBAL_STREAM.drawn
    assert np.linalg.norm(zeta - z_ref) <= 1e-6 * np.linalg.norm(z_ref)
    Jk = pd["bk1"] * sp.identity(m + n) + H_ref / pd["tk"]
    assert np.linalg.norm(Jk @ zeta - pd["z"]) <= 1e-9 * np.linalg.norm(pd["z"])



@pytest.mark.parametrize("tag", ["k30_s1", "k80_s2"])
def test_full_size_systems_three_solve_kernels(gpu, tag):
    """The same full-size systems through the three one-kernel forms of Class_AMG's solve loop -- cluster with the
    vectors in distributed shared memory (default), cluster with the vectors in global memory, grid-wide cooperative
    kernel: same cycle counts and component counts, solutions equal to 1e-9."""
    import os
    import torch
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "ssn_states_g128.npz"))
    g = int(d["g"]); m = n = g * g
    s = torch.zeros(m * n, dtype=torch.uint8, device="cuda"); s[torch.from_numpy(d[tag + "_lin"]).cuda()] = 1
    one = torch.ones(m, dtype=torch.float64, device="cuda")
    H = gpu.ASAt(s, one, one)
    del s
    pd = {"bk1": float(d[tag + "_bk1"]), "tk": float(d[tag + "_tk"]), "p": np.ones(m), "q": np.ones(n), "T": None, "H0": H, "z": d[tag + "_z"]}
    out = {}
    try:
        for mode in (2, 1, 0):
            gpu.set_cluster_solve(mode)
            gpu.rng_reset()
            if mode == 2:
                gpu.profile(True)
            zeta, it, res, info = gpu.Hybrid_AMG(pd, AMG_OPTS)
            if mode == 2:
                prof = gpu.profile_dump(); gpu.profile(False)
                assert "solve.dsm_solve_kernel" in prof and "solve.cluster_solve_kernel" not in prof, prof
            out[mode] = (np.asarray(zeta.cpu() if hasattr(zeta, "cpu") else zeta).reshape(-1), it, res, list(info))
    finally:
        gpu.set_cluster_solve(2); gpu.profile(False)
    z0, it0, res0, info0 = out[0]
    for mode in (1, 2):
        z1, it1, res1, info1 = out[mode]
        print(f"{tag}: solve kernel {mode} vs grid-wide: {it1} cycles (grid-wide {it0}), residual {res1:.2e} ({res0:.2e}), "
              f"solution dev {np.linalg.norm(z1 - z0) / np.linalg.norm(z0):.1e}")
        assert it1 == it0 and info1 == info0
        # each solve stops at a relative residual of 1e-11: the solutions agree to that times the conditioning of the state
        assert np.linalg.norm(z1 - z0) <= 1e-7 * np.linalg.norm(z0)


@pytest.mark.parametrize("tag", ["k30_s1", "k80_s2"])
def test_full_size_twogrid_cluster_kernel_equals_kernel_by_kernel(gpu, tag):
    """Hybrid_twogrid (inner_solver = 5) on full-size late-phase systems: the whole iteration loop of twogrid_bigph in ONE
    launch of the cluster kernel -- the coarse correction PCG(Ac, rrc, {[] -> 1e-11, 100, Jacobi}) of AMG/twogrid_bigph.m:98-99
    on vectors in distributed shared memory (default) -- against the kernel-by-kernel loop with the grid-wide pcg_kernel:
    same components, same number of two-grid iterations and final residual, solutions equal to 1e-7 (the conditioning of
    the late states); the profile shows which one ran."""
    import os
    import torch
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "ssn_states_g128.npz"))
    g = int(d["g"]); m = n = g * g
    s = torch.zeros(m * n, dtype=torch.uint8, device="cuda"); s[torch.from_numpy(d[tag + "_lin"]).cuda()] = 1
    one = torch.ones(m, dtype=torch.float64, device="cuda")
    H = gpu.ASAt(s, one, one)
    del s
    pd = {"bk1": float(d[tag + "_bk1"]), "tk": float(d[tag + "_tk"]), "p": np.ones(m), "q": np.ones(n), "T": None, "H0": H, "z": d[tag + "_z"]}
    out = {}
    try:
        for mode in (2, 0):
            gpu.set_cluster_solve(mode)
            gpu.rng_reset()
            gpu.profile(True)
            zeta, it, res, info = gpu.Hybrid_twogrid(pd, AMG_OPTS)
            prof = gpu.profile_dump(); gpu.profile(False)
            assert ("solve.dsm_solve_kernel" in prof) == (mode == 2), prof
            out[mode] = (np.asarray(zeta.cpu() if hasattr(zeta, "cpu") else zeta).reshape(-1), it, res, list(info))
    finally:
        gpu.set_cluster_solve(2); gpu.profile(False)
    (z1, it1, res1, info1), (z0, it0, res0, info0) = out[2], out[0]
    print(f"{tag}: two-grid in the cluster kernel: {it1} iterations (kernel by kernel {it0}), residual {res1:.2e} ({res0:.2e}), "
          f"solution dev {np.linalg.norm(z1 - z0) / np.linalg.norm(z0):.1e}")
    assert it1 == it0 and info1 == info0 and abs(res1 - res0) <= 1e-3 * res0
    # each solve stops at a relative residual of ~5e-12: the solutions agree to that times the conditioning of the state
    # (measured: 1.3e-14 at k30_s1, 2.7e-8 at k80_s2, as the three kernels of the W-cycle solver do there)
    assert np.linalg.norm(z1 - z0) <= 1e-7 * np.linalg.norm(z0)


@pytest.mark.parametrize("dens,unit", [(0.25, True), (0.25, False), (0.02, False)])
def test_hybrid_twogrid_matches_oracle(gpu, oracle, dens, unit):
    """Hybrid_twogrid.m / AMG/twogrid_bigph.m (inner_solver = 5) through the C ABI against the oracle: same
    components, same number of two-grid iterations, solution <= 1e-7 relative, KKT residual <= 1e-8."""
    import scipy.sparse as sp
    rs = np.random.RandomState(14)
    m, n = 150, 130
    s = rs.random_sample(m * n) < dens
    p, q = (np.ones(m), np.ones(n)) if unit else (rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5)
    H0 = oracle.ASAt(s, p, q)
    z = rs.standard_normal(m + n)
    opts = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 0.25, "smoth": 5, "cycle": "w", "isnsp": 1, "inter": 1, "guess": None}
    pd = {"bk1": 0.3, "tk": 0.8, "p": p, "q": q, "T": sp.diags(np.zeros(m + n)), "H0": H0, "z": z}
    oracle.rng_reset(); gpu.rng_reset()
    z_ref, it_ref, res_ref, info_ref = oracle.Hybrid_twogrid(pd, opts)
    zeta, it, res, info = gpu.Hybrid_twogrid(pd, opts)
    assert list(info) == list(info_ref)
    assert it == it_ref
    assert np.linalg.norm(zeta - z_ref) <= 1e-7 * np.linalg.norm(z_ref)
    Jk = 0.3 * sp.identity(m + n) + H0 / 0.8
    assert np.linalg.norm(Jk @ zeta - z) <= 1e-8 * np.linalg.norm(z)


def test_twogrid_bigph_matches_oracle(gpu, oracle):
    """twogrid_bigph on its own, with explicit options and with the function's own defaults for empty fields
    (retol [] -> 0, i.e. run to maxit): iteration counts and residual histories as the oracle's."""
    import scipy.sparse as sp
    rs = np.random.RandomState(3)
    m, n = 120, 100
    s = rs.random_sample(m * n) < 0.2
    H0 = oracle.ASAt(s, np.ones(m), np.ones(n))
    A = (0.1 * sp.identity(m + n) + H0).tocsc()
    A = (sp.diags(A.diagonal()) - (A - sp.diags(A.diagonal()))).tocsc()          # [V -U; -U' T]
    b = rs.standard_normal(m + n)
    for o in ({"retol": 1e-10, "maxit": 40, "fnode": n, "smoth": 3, "isnsp": 0, "guess": None},
              {"retol": None, "maxit": 4, "fnode": n, "smoth": None, "isnsp": None, "guess": None}):
        x_ref, it_ref, rel_ref, relk_ref, rhok_ref = oracle.twogrid_bigph(A, b, o)
        x, it, rel, relk, rhok = gpu.twogrid_bigph(A, b, o)
        assert it == it_ref and len(relk) == len(relk_ref)
        assert np.allclose(relk, relk_ref, rtol=1e-5, atol=1e-13)
        assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_ref)


def test_generic_twogrid_matches_oracle(gpu, oracle):
    """AMG/twogrid.m: bigph = 0 (damped Jacobi, mis_set(A,1/4) + standard interpolation; consumes the random
    stream) on a shifted grid Laplacian and bigph = 1 (the twogrid_bigph path) on a bigraph system, both
    against the oracle: iteration counts, histories, solutions."""
    import scipy.sparse as sp
    g = 20
    T = sp.diags([-np.ones(g - 1), 2 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1])
    A = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g)) + 1e-3 * sp.identity(g * g)).tocsc()
    b = np.random.RandomState(1).standard_normal(g * g)
    o = {"retol": 1e-10, "bigph": 0, "maxit": 60, "smoth": 3, "isnsp": 1, "guess": None}
    oracle.rng_reset(); gpu.rng_reset()
    x_ref, it_ref, rel_ref, relk_ref, _ = oracle.twogrid(A, b, o)
    x, it, rel, relk, _ = gpu.twogrid(A, b, o)
    assert it == it_ref and np.allclose(relk, relk_ref, rtol=1e-5, atol=1e-13)
    assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_ref)
    rs = np.random.RandomState(3)
    m, n = 90, 70
    s = rs.random_sample(m * n) < 0.2
    H0 = oracle.ASAt(s, np.ones(m), np.ones(n))
    B = (0.1 * sp.identity(m + n) + H0).tocsc()
    B = (sp.diags(B.diagonal()) - (B - sp.diags(B.diagonal()))).tocsc()
    b2 = rs.standard_normal(m + n)
    o2 = {"retol": 1e-10, "bigph": 1, "maxit": 40, "fnode": n, "smoth": 3, "isnsp": 0, "guess": None}
    x_ref, it_ref, _, relk_ref, _ = oracle.twogrid(B, b2, o2)
    x, it, _, relk, _ = gpu.twogrid(B, b2, o2)
    assert it == it_ref and np.allclose(relk, relk_ref, rtol=1e-5, atol=1e-13)
    assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_ref)
    with pytest.raises(Exception):
        gpu.twogrid(B, b2, {"retol": 1e-10, "bigph": 1, "maxit": 4, "smoth": 3, "isnsp": 0, "guess": None})   # fnode missing


@pytest.mark.parametrize("precd", [3, 4])
def test_pcg_ssor_and_ichol_match_oracle(gpu, oracle, precd):
    """PCG.m precd 3 (SSOR) and 4 (ichol) through the C ABI (level-scheduled sparse triangular solves inside
    the persistent PCG kernel) against the oracle: a grid Laplacian (many dependency levels) and a KKT matrix
    Jk = bk1*I + H0/tk of the path (two levels); iteration counts, residual histories, solutions."""
    import scipy.sparse as sp
    g = 24
    T = sp.diags([-np.ones(g - 1), 2 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1])
    A = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g)) + 0.05 * sp.identity(g * g)).tocsc()
    rs = np.random.RandomState(6)
    m, n = 130, 110
    s = rs.random_sample(m * n) < 0.05
    H0 = oracle.ASAt(s, rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5)
    Jk = (0.3 * sp.identity(m + n) + H0 / 0.8).tocsc()
    for M in (A, Jk):
        b = rs.standard_normal(M.shape[0])
        o = {"retol": 1e-11, "maxit": 2000, "precd": precd, "guess": None}
        d_ref, it_ref, res_ref, resk_ref = oracle.PCG(M, b, o)
        d, it, res, resk = gpu.PCG(M, b, o)
        assert abs(it - it_ref) <= 1, (it, it_ref)
        k = min(it, it_ref) - 1
        assert np.allclose(resk[:k], resk_ref[:k], rtol=1e-5, atol=1e-14)
        assert np.linalg.norm(d - d_ref) <= 1e-8 * np.linalg.norm(d_ref)
        assert np.linalg.norm(M @ d - b) <= 1e-9 * np.linalg.norm(b)
    if precd == 4:
        with pytest.raises(Exception):
            gpu.PCG((A - 10 * sp.identity(g * g)).tocsc(), rs.standard_normal(g * g), {"retol": 1e-11, "maxit": 10, "precd": 4, "guess": None})
