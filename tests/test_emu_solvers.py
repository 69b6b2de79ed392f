"""CPU-only: the REAL text of csrc/solvers.cu -- ASAt assembly (pattern and values bit for bit, also from the sorted
linear indices the row-sharded path exchanges), ASAtz, the rescaled system of Hybrid_AMG.m:17-24, components, invAAt --
compiled with g++ against tests/emu/common.cuh and compared with the oracle, as tests/test_gpu_plan.py and
tests/test_gpu_solvers.py do on the device.  (The active-set compaction kernels live in plan_ops.cu, which is not
emulated: the harness restates their contract on the host.)"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import random_active_problem
from test_emu_amg import _check, _csr_args, _fetch, _p, assert_same_matrix


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    import emu_build
    lib = emu_build.build(tmp_path_factory.mktemp("emu_solvers"), "emu_solvers.cpp",
                          ["solvers.cu", "sparse.cu", "amg_setup.cu", "amg_setup_fused.cu", "amg.cuh", "sparse.cuh", "solvers.cuh", "plan_ops.cuh"], "libemu_solvers.so")
    lib.emu_error.restype = C.c_char_p
    lib.emu_host_reads.restype = C.c_int64
    return lib


def asat(lib, s, p, q):
    m, n = p.size, q.size
    s8 = np.ascontiguousarray(s, dtype=np.uint8)
    _check(lib, lib.emu_asat(_p(s8), _p(p), _p(q), C.c_int64(m), C.c_int64(n)))
    return _fetch(lib)


@pytest.mark.parametrize("m,n,density,unit", [pytest.param(40, 24, 0.1, True, marks=__import__("emu_build").slow), (33, 47, 0.05, False),
                                                pytest.param(64, 16, 0.3, False, marks=__import__("emu_build").slow)])
def test_asat_pattern_and_values_exact(emu, oracle, m, n, density, unit):
    s, p, q = random_active_problem(m, n, density, seed=m + n, weights=not unit)
    H_ref = oracle.ASAt(s, p, q)
    assert_same_matrix(asat(emu, s, p, q), H_ref, "ASAt")
    # the same matrix from the sorted global linear indices (find(s)): the row-sharded assembly
    lin = np.flatnonzero(s).astype(np.int64)
    r0 = emu.emu_host_reads()
    _check(emu, emu.emu_asat_coo(_p(lin), C.c_int64(lin.size), _p(p), _p(q), C.c_int64(m), C.c_int64(n)))
    assert emu.emu_host_reads() - r0 == 1                             # one device->host read (nnz of H) per assembly
    assert_same_matrix(_fetch(emu), H_ref, "ASAt from linear indices")


def test_asat_edge_cases(emu, oracle):
    m, n = 40, 24
    p, q = np.ones(m), np.ones(n)
    H = asat(emu, np.zeros(m * n, dtype=bool), p, q)
    assert H.nnz == 0 and H.shape == (m + n, m + n)
    s = np.ones(m * n, dtype=bool)
    assert_same_matrix(asat(emu, s, p, q), oracle.ASAt(s, p, q), "ASAt, full active set")
    S = np.zeros((m, n), dtype=bool); S[3, :] = True; S[:, 5] = True              # isolated rows / columns elsewhere
    s = S.reshape(-1, order="F")
    assert_same_matrix(asat(emu, s, p, q), oracle.ASAt(s, p, q), "ASAt, one row and one column")


def test_asatz(emu, oracle):
    m = n = 48
    s, p, q = random_active_problem(m, n, 0.1, seed=9, weights=True)
    z = np.random.RandomState(1).standard_normal(m + n)
    y = np.zeros(m + n); s8 = s.astype(np.uint8)
    _check(emu, emu.emu_asatz(_p(z), _p(s8), _p(p), _p(q), C.c_int64(m), C.c_int64(n), _p(y)))
    assert np.allclose(y, oracle.ASAtz(z, s, p, q), rtol=1e-12, atol=1e-12)
    st = emu.emu_asatz(_p(np.zeros(7)), _p(np.zeros(12, np.uint8)), _p(np.ones(3)), _p(np.ones(4)), C.c_int64(3), C.c_int64(4), _p(np.zeros(7)))
    assert st == -14                                                  # SSN_E_ASATZ_DIM


def components(lib, A):
    A, a = _csr_args(A)
    n = A.shape[0]
    blocks, sizes, perm, r = (np.zeros(n + 1, np.int32) for _ in range(4))
    nc = C.c_int(0)
    st = lib.emu_components(C.c_int64(n), C.c_int64(A.shape[1]), C.c_int64(A.nnz), _p(a[0]), _p(a[1]), _p(a[2]), _p(blocks), _p(sizes), _p(perm), _p(r), C.byref(nc))
    return st, blocks[:n], sizes[:nc.value], perm[:n], r[:nc.value + 1]


def test_components_ordering(emu, oracle):
    for seed, n, d in [(0, 60, 0.02), (2, 150, 0.01)]:
        rs = np.random.RandomState(seed)
        A = sp.random(n, n, density=d, random_state=rs, format="csr"); A = (A + A.T + sp.identity(n)).tocsr()
        b_ref, s_ref, p_ref, r_ref = oracle.components(A)
        st, b, s, p, r = components(emu, A)
        assert st == 0
        assert np.array_equal(b, b_ref) and np.array_equal(s, s_ref)  # 1-based labels, 0-based members / boundaries on both sides
        assert np.array_equal(p, p_ref) and np.array_equal(r, r_ref)
    n = 300                                                           # a long path: pointer jumping must converge
    P = sp.diags([np.ones(n - 1), np.ones(n - 1), 2 * np.ones(n)], [-1, 1, 0], format="csr")
    st, b, s, p, r = components(emu, P)
    assert st == 0 and s.tolist() == [n] and np.array_equal(p, np.arange(n))
    st = components(emu, sp.random(4, 5, density=0.5, format="csr", random_state=1))[0]
    assert st == -6                                                   # SSN_E_NOT_SQUARE


@pytest.mark.parametrize("weights", [False, True])
def test_rescaled_system_bit_exact(emu, oracle, weights):
    from oracle.solvers import rescaled_system
    m, n = 50, 40
    s, p, q = random_active_problem(m, n, 0.06, 11, weights)
    t = np.random.RandomState(0).random_sample(m + n) * (np.arange(m + n) % 3 == 0)
    H0 = oracle.ASAt(s, p, q)
    z = np.random.RandomState(11).standard_normal(m + n)
    pd = {"bk1": 0.05, "tk": 0.8, "p": p, "q": q, "T": sp.diags(t), "H0": H0, "z": z}
    qp, A0, Qd, Kd, Ae_ref, f_ref = rescaled_system(pd)
    H, h = _csr_args(H0)
    f = np.zeros(m + n)
    _check(emu, emu.emu_rescaled_system(C.c_double(0.05), C.c_double(0.8), C.c_int64(m), C.c_int64(n), _p(p), _p(q), _p(t), _p(z),
                                        C.c_int64(H.nnz), _p(h[0]), _p(h[1]), _p(h[2]), _p(f)))
    assert_same_matrix(_fetch(emu), Ae_ref, "Ae")
    assert np.array_equal(f, f_ref)
    pz = p.copy(); pz[3] = 0.0
    st = emu.emu_rescaled_system(C.c_double(0.05), C.c_double(0.8), C.c_int64(m), C.c_int64(n), _p(pz), _p(q), _p(t), _p(z),
                                 C.c_int64(H.nnz), _p(h[0]), _p(h[1]), _p(h[2]), _p(f))
    assert st == -3 and b"p or q contains 0" in emu.emu_error()       # Hybrid_AMG.m:18-19


def test_invaat(emu, oracle):
    m, n = 31, 22
    rs = np.random.RandomState(3)
    p, q = rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5
    x = rs.standard_normal(m + n)
    for sg1, sg2 in [(0.7, 0.7), (0.7, 1.9)]:
        y = np.zeros(m + n)
        _check(emu, emu.emu_invaat(_p(x), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(sg1), C.c_double(sg2), _p(y)))
        assert np.allclose(y, oracle.invAAt(x, p, q, sg1, sg2), rtol=1e-12, atol=1e-13)
