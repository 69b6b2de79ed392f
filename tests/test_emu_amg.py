"""CPU-only: the REAL text of csrc/sparse.cu and csrc/amg_setup.cu (SpGEMM in the frozen order, transpose, sparse add,
strength, the MT19937 stream, MIS / C-F splitting, interpolation and the Galerkin products of transfer) compiled with
g++ against tests/emu/common.cuh -- one host thread per CUDA thread, block by block, meeting at __syncthreads and at
the warp votes / shuffles -- and compared with the oracle bit for bit, as tests/test_gpu_amg.py does on the device.
Small systems only (every CUDA thread is a host thread)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import random_active_problem


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    import emu_build
    lib = emu_build.build(tmp_path_factory.mktemp("emu_amg"), "emu_amg.cpp", ["sparse.cu", "amg_setup.cu", "amg_setup_fused.cu", "amg_cluster.cu", "amg.cuh", "sparse.cuh"],
                          "libemu_amg.so")
    lib.emu_error.restype = C.c_char_p
    lib.emu_rng_drawn.restype = C.c_int64
    lib.emu_launches.restype = C.c_int64
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _csr_args(A):
    A = sp.csr_matrix(A); A.sort_indices()
    ptr, idx, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    return A, (ptr, idx, val)


def _check(lib, st):
    assert st == 0, f"{st}: {lib.emu_error().decode()}"


def _fetch(lib, k=0):
    sz = np.zeros(3, np.int64); lib.emu_sizes(C.c_int(k), _p(sz))
    nr, nc, nnz = (int(v) for v in sz)
    ptr = np.zeros(nr + 1, np.int32); idx = np.zeros(max(nnz, 1), np.int32); val = np.zeros(max(nnz, 1))
    lib.emu_fetch(C.c_int(k), _p(ptr), _p(idx), _p(val))
    return sp.csr_matrix((val[:nnz], idx[:nnz], ptr), shape=(nr, nc))


def _flags(lib, k, n):
    out = np.zeros(n, np.uint8); lib.emu_fetch_flags(C.c_int(k), _p(out)); return out.astype(bool)


def assert_same_matrix(D, R, what):
    """pattern exact, values bit for bit (both sides in CSC with sorted indices, like tests/test_gpu_amg.py)"""
    D = sp.csc_matrix(D); R = sp.csc_matrix(R); D.sort_indices(); R.sort_indices()
    assert D.shape == R.shape, what
    assert np.array_equal(D.indptr, R.indptr) and np.array_equal(D.indices, R.indices), what + ": pattern differs"
    bad = np.flatnonzero(D.data != R.data)
    assert bad.size == 0, f"{what}: {bad.size} values differ, first {D.data[bad[:3]]} vs {R.data[bad[:3]]}"


def rand_sparse(nr, nc, density, seed):
    rs = np.random.RandomState(seed)
    A = sp.random(nr, nc, density=density, random_state=rs, format="csr", data_rvs=rs.standard_normal)
    A.sort_indices()
    return A


def spgemm(lib, A, B):
    A, a = _csr_args(A); B, b = _csr_args(B)
    _check(lib, lib.emu_spgemm(C.c_int64(A.shape[0]), C.c_int64(A.shape[1]), C.c_int64(A.nnz), _p(a[0]), _p(a[1]), _p(a[2]),
                               C.c_int64(B.shape[0]), C.c_int64(B.shape[1]), C.c_int64(B.nnz), _p(b[0]), _p(b[1]), _p(b[2])))
    return _fetch(lib)


def ssn_matrix(oracle, m, n, density, seed, weights=False, bk1=0.05, tk=0.8):
    s, p, q = random_active_problem(m, n, density, seed, weights)
    H0 = oracle.ASAt(s, p, q)
    pd = {"bk1": bk1, "tk": tk, "p": p, "q": q, "T": sp.diags(np.zeros(m + n)), "H0": H0, "z": np.zeros(m + n)}
    from oracle.solvers import rescaled_system
    return rescaled_system(pd)[4]


def test_mt19937_stream(emu, oracle):
    """MATLAB's default stream (mt19937ar, seed 5489, 53-bit doubles): two draws that cross a 624-word state refill."""
    oracle.rng_reset()
    _check(emu, emu.emu_rng_reset())
    for count in (5, 700, 33):
        out = np.zeros(count)
        _check(emu, emu.emu_rand(C.c_int64(count), _p(out)))
        assert np.array_equal(out, oracle.rand(count))
    assert emu.emu_rng_drawn() == 738


@pytest.mark.parametrize("shape", [(30, 40, 50, 0.2), (120, 90, 150, 0.05), (20, 600, 700, 0.1)])
def test_spgemm_fixed_order_bit_exact(emu, oracle, shape):
    from oracle.amg import spgemm as ref
    nr, nk, nc, d = shape
    A = rand_sparse(nr, nk, d, 1); B = rand_sparse(nk, nc, d, 2)
    assert_same_matrix(spgemm(emu, A, B), ref(A, B), "A*B")


def test_spgemm_identity_runs_and_cancellation(emu, oracle):
    from oracle.amg import spgemm as ref
    n = 60
    P = sp.vstack([rand_sparse(50, n, 0.08, 3), sp.identity(n, format="csr")]).tocsr()     # [W ; I]
    A = rand_sparse(110, 110, 0.06, 4); A = (A + A.T).tocsr()
    T1 = ref(P.T, A)
    assert_same_matrix(spgemm(emu, P.T.tocsr(), A), T1, "P'*A")
    assert_same_matrix(spgemm(emu, T1, P), ref(T1, P), "(P'*A)*P")
    X = sp.csr_matrix(np.array([[1.0, -1.0], [2.0, 3.0]])); Y = sp.csr_matrix(np.array([[1.0, 5.0], [1.0, 7.0]]))
    assert_same_matrix(spgemm(emu, X, Y), ref(X, Y), "exact zero dropped")


def test_transpose_and_sparse_add(emu):
    A, a = _csr_args(rand_sparse(70, 130, 0.05, 5))
    _check(emu, emu.emu_transpose(C.c_int64(70), C.c_int64(130), C.c_int64(A.nnz), _p(a[0]), _p(a[1]), _p(a[2])))
    assert_same_matrix(_fetch(emu), A.T, "A'")
    B, b = _csr_args(rand_sparse(70, 130, 0.05, 6))
    _check(emu, emu.emu_sparse_add(C.c_int64(70), C.c_int64(130), C.c_int64(A.nnz), _p(a[0]), _p(a[1]), _p(a[2]), C.c_double(0.5),
                                   C.c_int64(B.nnz), _p(b[0]), _p(b[1]), _p(b[2])))
    R = (A + 0.5 * B).tocsr(); R.eliminate_zeros()
    assert_same_matrix(_fetch(emu), R, "A + 0.5*B")


@pytest.mark.parametrize("m,n,density,weights", [(40, 30, 0.1, False), pytest.param(70, 60, 0.05, True, marks=__import__("emu_build").slow)])
def test_strength_mis_set_and_transfer_bit_exact(emu, oracle, m, n, density, weights):
    """The sequence of tests/test_gpu_amg.py::test_mis_set_and_transfer_bit_exact on the emulated sources: level 1 (bigraph
    branch, transfer.m:19-29), strength of both levels, level 2 through mis_set (same random stream) and the standard
    interpolation with and without the row normalisation."""
    from oracle.amg import transfer
    Ae = ssn_matrix(oracle, m, n, density, seed=m + 1, weights=weights)
    N = m + n
    o = {"theta": 0.25, "bigph": 1, "inter": 1, "isnsp": 1, "fnode": n}
    A1, a1 = _csr_args(Ae)

    def run_transfer(A, a, bigph, isnsp, J):
        _check(emu, emu.emu_transfer(C.c_int64(A.shape[0]), C.c_int64(A.nnz), _p(a[0]), _p(a[1]), _p(a[2]), C.c_double(0.25), C.c_int(bigph),
                                     C.c_int(1), C.c_int(isnsp), C.c_int(n), C.c_int(J)))
        return _fetch(emu, 0), _fetch(emu, 1), _fetch(emu, 2), _flags(emu, 0, A.shape[0])

    Ac_ref, Pro_ref, As_ref, indC_ref = transfer(Ae, o, J=1, want_aux=True)
    Ac, Pro, As, indC = run_transfer(A1, a1, 1, 1, 1)
    assert np.array_equal(indC, np.asarray(indC_ref).astype(bool).ravel())
    assert_same_matrix(Pro, Pro_ref, "Pro level 1")
    assert_same_matrix(As, As_ref, "As level 1")
    assert_same_matrix(Ac, Ac_ref, "Ac level 2")

    A2, a2 = _csr_args(Ac_ref)
    for A, a in ((A1, a1), (A2, a2)):
        for which in (1, 2):
            _check(emu, emu.emu_strength(C.c_int64(A.shape[0]), C.c_int64(A.nnz), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(which)))
            assert_same_matrix(_fetch(emu), oracle.strength(A, which), f"strength which={which}")

    oracle.rng_reset(); _check(emu, emu.emu_rng_reset())
    isC_ref, isF_ref, As2_ref = oracle.mis_set(Ac_ref, 0.25)
    _check(emu, emu.emu_mis_set(C.c_int64(A2.shape[0]), C.c_int64(A2.nnz), _p(a2[0]), _p(a2[1]), _p(a2[2]), C.c_double(0.25)))
    assert np.array_equal(_flags(emu, 0, A2.shape[0]), np.asarray(isC_ref).astype(bool).ravel())
    assert np.array_equal(_flags(emu, 1, A2.shape[0]), np.asarray(isF_ref).astype(bool).ravel())
    assert_same_matrix(_fetch(emu), As2_ref, "As level 2")
    assert emu.emu_rng_drawn() == oracle.GLOBAL_STREAM.drawn

    for isnsp in (1, 0):
        oracle.rng_reset(); _check(emu, emu.emu_rng_reset())
        A3_ref, P3_ref, As3_ref, indC3_ref = transfer(Ac_ref, dict(o, isnsp=isnsp), J=2, want_aux=True)
        A3, P3, As3, indC3 = run_transfer(A2, a2, 1, isnsp, 2)
        assert np.array_equal(indC3, np.asarray(indC3_ref).astype(bool).ravel())
        assert_same_matrix(P3, P3_ref, f"Pro level 2 isnsp={isnsp}")
        assert_same_matrix(A3, A3_ref, f"Ac level 3 isnsp={isnsp}")


def test_hierarchy_bit_exact(emu, oracle):
    """amg_setup (Class_AMG.m:41-85) on the emulated sources: level sizes, every A_k and Pro_k bit for bit, the random
    draws consumed, and ones'*A_k*ones of every level (read back once for the whole hierarchy)."""
    from oracle.amg import setup_hierarchy, amg_state
    m, n = 60, 50
    Ae = ssn_matrix(oracle, m, n, 0.06, seed=3 * m)
    if oracle.components(Ae)[1].size != 1:
        pytest.skip("random active set is disconnected")
    o = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 0.25, "smoth": 5, "cycle": "w", "isnsp": 1, "inter": 1, "guess": None, "fnode": n}
    oracle.rng_reset(); _check(emu, emu.emu_rng_reset())
    st = setup_hierarchy(Ae, o)
    ref = [(A.copy(), P) for A, P in zip(st.Ack, st.Prok)]
    A1, a1 = _csr_args(Ae)
    J = C.c_int(0)
    _check(emu, emu.emu_amg_setup(C.c_int64(m + n), C.c_int64(A1.nnz), _p(a1[0]), _p(a1[1]), _p(a1[2]), C.c_double(0.25), C.c_int(5), C.c_int(1),
                                  C.c_int(n), C.byref(J)))
    assert J.value == len(ref) and J.value >= 3
    emu.emu_level.restype = C.c_double
    for k, (A, P) in enumerate(ref):
        xx = emu.emu_level(C.c_int(k), C.c_int(0))
        assert_same_matrix(_fetch(emu), A, f"A level {k + 1}")
        assert abs(xx - A.sum()) <= 1e-12 * abs(A).sum(), f"ones'*A*ones level {k + 1}"
        if k > 0:
            emu.emu_level(C.c_int(k), C.c_int(1))
            assert_same_matrix(_fetch(emu), P, f"Pro level {k + 1}")
    assert emu.emu_rng_drawn() == oracle.GLOBAL_STREAM.drawn
    emu.emu_phase_counts.restype = C.c_char_p
    assert b"setup.fused_small_levels" in emu.emu_phase_counts(), "the levels below level 2 were not built by the fused kernel"
    amg_state.clear(); emu.emu_amg_clear()


@pytest.mark.parametrize("m,n,density,isnsp", [(90, 70, 0.05, 0), pytest.param(400, 300, 0.01, 1, marks=__import__("emu_build").slow)])
def test_fused_small_levels_equal_the_oracle(emu, oracle, m, n, density, isnsp):
    """amg_setup_fused.cu (all levels below level 2 in ONE kernel: strength, MIS rounds, interpolation, transposes, the two
    sparse products in their frozen summation order, smoother data) against the oracle's hierarchy, bit for bit."""
    from oracle.amg import setup_hierarchy, amg_state
    Ae = ssn_matrix(oracle, m, n, density, seed=7 * m)
    if oracle.components(Ae)[1].size != 1:
        pytest.skip("random active set is disconnected")
    o = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 0.25, "smoth": 5, "cycle": "w", "isnsp": isnsp, "inter": 1, "guess": None, "fnode": n}
    oracle.rng_reset(); _check(emu, emu.emu_rng_reset())
    st = setup_hierarchy(Ae, o)
    ref = [(A.copy(), P) for A, P in zip(st.Ack, st.Prok)]
    A1, a1 = _csr_args(Ae)
    J = C.c_int(0)
    emu.emu_phase_reset()
    _check(emu, emu.emu_amg_setup(C.c_int64(m + n), C.c_int64(A1.nnz), _p(a1[0]), _p(a1[1]), _p(a1[2]), C.c_double(0.25), C.c_int(5), C.c_int(isnsp),
                                  C.c_int(n), C.byref(J)))
    assert J.value == len(ref) and J.value >= 3
    emu.emu_level.restype = C.c_double
    for k, (A, P) in enumerate(ref):
        emu.emu_level(C.c_int(k), C.c_int(0))
        assert_same_matrix(_fetch(emu), A, f"A level {k + 1}")
        if k > 0:
            emu.emu_level(C.c_int(k), C.c_int(1))
            assert_same_matrix(_fetch(emu), P, f"Pro level {k + 1}")
    assert emu.emu_rng_drawn() == oracle.GLOBAL_STREAM.drawn
    emu.emu_phase_counts.restype = C.c_char_p
    counts = dict((ln.split("\t")[0], ln.split("\t")[1:]) for ln in emu.emu_phase_counts().decode().splitlines())
    # the first MIS level (the one the fused kernel starts from) is still coarsened piece by piece
    assert "setup.fused_small_levels" in counts, counts.keys()
    amg_state.clear(); emu.emu_amg_clear()


def test_scan_paths_agree(emu, oracle):
    """scan_counts_to_ptr through the one-block kernel and through cub::DeviceScan (what arrays above
    ssn_ctx::small_scan_max take): the same product either way."""
    from oracle.amg import spgemm as ref
    A = rand_sparse(120, 90, 0.05, 1); B = rand_sparse(90, 150, 0.05, 2)
    try:
        emu.emu_set_small_scan_max(C.c_int(0))
        assert_same_matrix(spgemm(emu, A, B), ref(A, B), "A*B, cub scans")
    finally:
        emu.emu_set_small_scan_max(C.c_int(1 << 14))
    assert_same_matrix(spgemm(emu, A, B), ref(A, B), "A*B, one-block scans")


def test_spgemm_in_slabs_of_rows(emu, oracle):
    """A product whose intermediate upper bound exceeds ssn_ctx::spgemm_slab_limit (2^30 entries in production: the early
    SsN steps of the 256x256 grids) is formed slab of rows by slab of rows: same matrix bit for bit."""
    from oracle.amg import spgemm as ref
    A = rand_sparse(120, 90, 0.3, 1); B = rand_sparse(90, 150, 0.3, 2)
    want = ref(A, B)
    try:
        for limit in (4000, 700, 1):                       # a few slabs, many slabs, one row per slab
            emu.emu_set_spgemm_slab(C.c_int64(limit))
            assert_same_matrix(spgemm(emu, A, B), want, f"A*B in slabs (limit {limit})")
    finally:
        emu.emu_set_spgemm_slab(C.c_int64(1 << 30))


_slow = __import__("emu_build").slow
@pytest.mark.parametrize("m,n,density,isnsp,cycle,bigph,noreg", [(90, 70, 0.05, 1, "w", 1, 0), pytest.param(90, 70, 0.05, 1, "w", 1, 1, marks=_slow),
                                                                  pytest.param(90, 70, 0.05, 0, "v", 1, 0, marks=_slow),
                                                                  pytest.param(60, 50, 0.08, 1, "w", 0, 0, marks=_slow), (60, 50, 0.08, 1, "w", 0, 1)])
def test_dsm_cluster_solve_against_the_oracle(emu, oracle, monkeypatch, m, n, density, isnsp, cycle, bigph, noreg):
    """amg_cluster.cu (Class_AMG's solve loop inside one cluster, level vectors in distributed shared memory, the
    two-half-sweep form of the bigraph smoother, op list of the cycle) on a cluster of 16 emulated CTAs: cycle counts
    and residual histories of the oracle's Class_AMG (AMG/Class_AMG.m:89-107), the solution to 1e-10."""
    from oracle.amg import setup_hierarchy, amg_state, MG_Wcycle, MG_Vcycle, Class_AMG
    monkeypatch.setenv("SSN_DSM_NOREG", str(noreg))     # 1: the smoothing loops re-read their rows every sweep (rows that do not fit the registers)
    halo = 1 if (m, noreg) in ((90, 0), (60, 1)) else 0  # the opt-in halo copies of the A-gathers in two of the cases
    monkeypatch.setenv("SSN_DSM_HALO", str(halo))
    Ae = ssn_matrix(oracle, m, n, density, seed=7 * m)
    if oracle.components(Ae)[1].size != 1:
        pytest.skip("random active set is disconnected")
    o = {"retol": 1e-11, "bigph": bigph, "maxit": 30, "theta": 0.25, "smoth": 3, "cycle": cycle, "isnsp": isnsp, "inter": 1, "guess": None, "fnode": n}
    rs = np.random.RandomState(5)
    b = rs.standard_normal(m + n)
    if isnsp:
        b -= b.mean()
    guess = 0.01 * rs.standard_normal(m + n)
    oracle.rng_reset()
    x_ref, it_ref, rel_ref, relk_ref, rho_ref = Class_AMG(Ae, b, dict(o, guess=guess))
    # the same hierarchy again, kept alive: the dense operator of level kd, column by column
    oracle.rng_reset(); _check(emu, emu.emu_rng_reset())
    st = setup_hierarchy(Ae, o)
    J = st.J
    assert J >= 3
    kd = J - 2 if J >= 4 else J - 1
    Nk = st.Ack[kd].shape[0]
    cyc = MG_Wcycle if cycle == "w" else MG_Vcycle
    B = np.column_stack([cyc(np.eye(Nk)[:, i], isnsp, kd + 1) for i in range(Nk)])
    A1, a1 = _csr_args(Ae)
    Jd = C.c_int(0)
    emu.emu_set_bigph(C.c_int(bigph))
    _check(emu, emu.emu_amg_setup(C.c_int64(m + n), C.c_int64(A1.nnz), _p(a1[0]), _p(a1[1]), _p(a1[2]), C.c_double(0.25), C.c_int(3), C.c_int(isnsp),
                                  C.c_int(n), C.byref(Jd)))
    emu.emu_set_bigph(C.c_int(1))
    assert Jd.value == J
    x = np.zeros(m + n); relk = np.zeros(40); rho = np.zeros(40)
    it = C.c_int(0); hl = C.c_int(0); status = C.c_int(-7)
    Bc = np.ascontiguousarray(B)
    _check(emu, emu.emu_dsm_solve(C.c_int(kd), _p(Bc), _p(b), _p(guess), C.c_int(isnsp), C.c_int(1 if cycle == "w" else 0), C.c_double(1e-11),
                                  C.c_int(30), _p(x), C.byref(it), _p(relk), _p(rho), C.byref(hl), C.byref(status)))
    amg_state.clear(); emu.emu_amg_clear()
    assert status.value == 0, status.value
    assert (emu.emu_last_halo() > 0) == (halo == 1)                     # the A-gathers of CTA 0 went through its halo copies
    assert it.value == it_ref and hl.value == len(relk_ref)
    big = relk_ref > 1e-10
    dev = np.max(np.abs(relk[:hl.value][big] - relk_ref[big]) / relk_ref[big])
    print(f"dsm cluster solve {m}x{n} isnsp={isnsp} {cycle}-cycle bigph={bigph} noreg={noreg}: {it.value} cycles, history dev {dev:.1e}, "
          f"solution {np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref):.1e}")
    assert dev <= 1e-6, dev                                             # the leaf is PCG to 1e-11 in the oracle, a dense operator here
    assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)


@pytest.mark.parametrize("m,n,density,isnsp,noreg", [(90, 70, 0.05, 1, 0), pytest.param(60, 50, 0.08, 0, 1, marks=_slow)])
def test_twogrid_cluster_kernel_against_the_oracle(emu, oracle, monkeypatch, m, n, density, isnsp, noreg):
    """twogrid_bigph's whole iteration loop inside the cluster kernel (amg_cluster.cu: Z_PCG, the coarse correction
    PCG(Ac, rrc, {[] -> 1e-11, 100, Jacobi}) of AMG/twogrid_bigph.m:98-99 on the level vectors in distributed shared memory)
    on 16 emulated CTAs against the oracle's twogrid_bigph: iteration count, residual history, solution."""
    monkeypatch.setenv("SSN_DSM_NOREG", str(noreg))
    monkeypatch.setenv("SSN_DSM_HALO", "0")
    monkeypatch.setenv("SSN_PCG_LT0", "1")               # one lane per row: the entries past the 8 in registers go through the shared-memory tail area
    Ae = ssn_matrix(oracle, m, n, density, seed=7 * m)
    rs = np.random.RandomState(11)
    b = rs.standard_normal(m + n)
    if isnsp:
        b -= b.mean()
    guess = 0.01 * rs.standard_normal(m + n)
    o = {"retol": 1e-11, "maxit": 30, "smoth": 3, "isnsp": isnsp, "guess": guess, "fnode": n}
    x_ref, it_ref, rel_ref, relk_ref, rho_ref = oracle.twogrid_bigph(Ae, b, o)
    A1, a1 = _csr_args(Ae)
    x = np.zeros(m + n); relk = np.zeros(40); rho = np.zeros(40)
    it = C.c_int(0); hl = C.c_int(0); status = C.c_int(-7)
    _check(emu, emu.emu_twogrid_dsm(C.c_int64(m + n), C.c_int64(A1.nnz), _p(a1[0]), _p(a1[1]), _p(a1[2]), C.c_int(3), C.c_int(isnsp), C.c_int(n),
                                    _p(b), _p(guess), C.c_double(1e-11), C.c_int(30), _p(x), C.byref(it), _p(relk), _p(rho), C.byref(hl), C.byref(status)))
    assert status.value == 0, status.value
    relk_ref = np.asarray(relk_ref, dtype=float).reshape(-1)
    assert it.value == it_ref and hl.value == len(relk_ref), (it.value, it_ref, hl.value, len(relk_ref))
    big = relk_ref > 1e-9
    assert np.allclose(relk[:hl.value][big], relk_ref[big], rtol=1e-6, atol=0)
    assert np.linalg.norm(x - x_ref) <= 1e-9 * max(1.0, np.linalg.norm(x_ref))
