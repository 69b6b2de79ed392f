"""CPU-only: the REAL text of csrc/trifactor.cu (device-side SSOR / IC(0) factor setup, dependency levels, Jk assembly)
compiled with g++ against tests/emu/common.cuh, a host stand-in for the CUDA names it uses, and run against the oracle.
Every thread of every launch is executed (one after the other; the 32 lanes of a warp as 32 host threads for the
kernels that vote).  This checks indexing, ordering and arithmetic of the source nvcc compiles -- not what only the
hardware can show; the GPU run of the same paths is tests/test_zz_device_setup.py."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT



@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    import emu_build
    return emu_build.build(tmp_path_factory.mktemp("emu"), "emu_trifactor.cpp", ["trifactor.cu", "sparse.cu", "amg.cuh", "sparse.cuh"], "libemu_trifactor.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def tri_factors(lib, H, precd):
    H = sp.csr_matrix(H); H.sort_indices()
    n, nnz = H.shape[0], H.nnz
    ptr, idx, val = H.indptr.astype(np.int32), H.indices.astype(np.int32), H.data.astype(np.float64)
    cap = nnz + n
    out = {k: np.full(cap, -7, np.int32) for k in ("li", "ui")}
    out.update({k: np.full(cap, np.nan) for k in ("lv", "uv")})
    out.update({k: np.full(n + 1, -7, np.int32) for k in ("lp", "up", "llev", "ulev")})
    out.update({k: np.full(n, -7, np.int32) for k in ("lrows", "urows")})
    out["mid"] = np.full(n, np.nan)
    sizes = np.zeros(6, np.int64); err = C.create_string_buffer(256)
    st = lib.emu_tri_factors(C.c_int(n), C.c_int64(nnz), _p(ptr), _p(idx), _p(val), C.c_int(precd), _p(sizes),
                             _p(out["lp"]), _p(out["li"]), _p(out["lv"]), _p(out["up"]), _p(out["ui"]), _p(out["uv"]), _p(out["mid"]),
                             _p(out["lrows"]), _p(out["llev"]), _p(out["urows"]), _p(out["ulev"]), err, C.c_int(256))
    if st != 0:
        raise RuntimeError(f"{st}: {err.value.decode()}")
    nl, nu, levl, levu = (int(v) for v in sizes[:4])
    L = sp.csr_matrix((out["lv"][:nl], out["li"][:nl], out["lp"]), shape=(n, n))
    U = sp.csr_matrix((out["uv"][:nu], out["ui"][:nu], out["up"]), shape=(n, n))
    return {"L": L, "U": U, "mid": out["mid"] if sizes[4] else None, "lrows": out["lrows"], "llev": out["llev"][:levl + 1],
            "urows": out["urows"], "ulev": out["ulev"][:levu + 1], "launches": int(sizes[5])}


def check_levels(T, rows, levptr, lower):
    """rows grouped by dependency level: a permutation, ascending inside a level, every row one level above the deepest
    row it depends on (= the longest-path levels a sequential sweep computes)."""
    n = T.shape[0]
    assert sorted(rows.tolist()) == list(range(n)) and levptr[0] == 0 and levptr[-1] == n
    lev = np.empty(n, np.int64)
    for l in range(len(levptr) - 1):
        seg = rows[levptr[l]:levptr[l + 1]]
        assert len(seg) > 0 and np.all(np.diff(seg) > 0)
        lev[seg] = l
    ref = np.zeros(n, np.int64)
    order = range(n) if lower else range(n - 1, -1, -1)
    for i in order:
        cols = T.indices[T.indptr[i]:T.indptr[i + 1]]
        deps = cols[cols < i] if lower else cols[cols > i]
        ref[i] = 0 if len(deps) == 0 else ref[deps].max() + 1
    assert np.array_equal(lev, ref)


def systems(oracle):
    g = 13
    T = sp.diags([-np.ones(g - 1), 2 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1])
    lap = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g)) + 0.05 * sp.identity(g * g)).tocsr()
    rs = np.random.RandomState(6)
    m, n = 60, 45
    S = rs.random_sample((m, n)) < 0.08
    S[7, :] = False                                                      # a row node without active entries
    H0 = oracle.ASAt(S.reshape(-1, order="F"), rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5)
    Jk = (0.3 * sp.identity(m + n) + H0 / 0.8).tocsr()
    B = sp.random(80, 80, density=0.06, random_state=3); R = (B + B.T + 12 * sp.identity(80)).tocsr()
    tri = sp.diags([-np.ones(49), 2.5 * np.ones(50), -np.ones(49)], [-1, 0, 1]).tocsr()
    return {"laplacian": lap, "Jk": Jk, "random_spd": R, "tridiagonal": tri}


@pytest.mark.parametrize("name", ["laplacian", "Jk", "random_spd", "tridiagonal"])
def test_ic0_factor_is_the_oracles_bit_for_bit(emu, oracle, name):
    from oracle.pcg import ichol0
    H = systems(oracle)[name]
    F = tri_factors(emu, H, 4)
    Lref = sp.csr_matrix(ichol0(H)); Lref.sort_indices()
    L = F["L"]
    assert np.array_equal(L.indptr, Lref.indptr) and np.array_equal(L.indices, Lref.indices)
    assert np.array_equal(L.data, Lref.data)                             # same summation order, no FMA
    for i in range(H.shape[0]):                                          # diagonal last in Lf, first in Uf
        assert L.indices[L.indptr[i + 1] - 1] == i and F["U"].indices[F["U"].indptr[i]] == i
    Ut = sp.csr_matrix(L.T); Ut.sort_indices()
    assert np.array_equal(F["U"].indptr, Ut.indptr) and np.array_equal(F["U"].indices, Ut.indices) and np.array_equal(F["U"].data, Ut.data)
    assert F["mid"] is None
    check_levels(L, F["lrows"], F["llev"], lower=True)
    check_levels(F["U"], F["urows"], F["ulev"], lower=False)
    if name == "Jk":
        assert len(F["llev"]) - 1 == 2 and len(F["ulev"]) - 1 == 2       # bipartite: column nodes, then row nodes
    if name == "tridiagonal":
        assert len(F["llev"]) - 1 == 50                                  # one row per level: the relaxation needs n sweeps
        assert abs(L @ L.T - H).max() < 1e-14                            # IC(0) of a tridiagonal matrix is exact


@pytest.mark.parametrize("name", ["laplacian", "Jk", "random_spd"])
def test_ssor_factors(emu, oracle, name):
    """PCG.m:39-44,96-99: Lf = D + w*L, mid = D, Uf = (w*(2-w))*(D + w*U), w = 1.5, one rounding per product."""
    H = systems(oracle)[name]
    if name == "random_spd":
        H = H.tolil(); H[5, 5] = 0; H = sp.csr_matrix(H); H.eliminate_zeros()          # a missing diagonal is a stored zero
    F = tri_factors(emu, H, 3)
    om = 1.5; sc = om * (2.0 - om)
    D = H.diagonal()
    Lref = (sp.tril(H, -1) * om).tocsr(); Uref = (sp.triu(H, 1) * om * 1.0).tocsr()
    Lfull = F["L"].toarray(); Ufull = F["U"].toarray()
    assert np.array_equal(np.tril(Lfull, -1), Lref.toarray()) and np.array_equal(np.diag(Lfull), D)
    assert np.array_equal(np.triu(Ufull, 1), sc * Uref.toarray()) and np.array_equal(np.diag(Ufull), sc * D)
    assert np.array_equal(F["mid"], D)
    n = H.shape[0]
    assert F["L"].nnz == sp.tril(H, -1).nnz + n and F["U"].nnz == sp.triu(H, 1).nnz + n
    for i in range(n):
        assert F["L"].indices[F["L"].indptr[i + 1] - 1] == i and F["U"].indices[F["U"].indptr[i]] == i
    check_levels(F["L"], F["lrows"], F["llev"], lower=True)
    check_levels(F["U"], F["urows"], F["ulev"], lower=False)


def test_ic0_error_statuses(emu, oracle):
    H = systems(oracle)["laplacian"]
    with pytest.raises(RuntimeError, match="-12: ichol: encountered nonpositive pivot"):
        tri_factors(emu, H - 10 * sp.identity(H.shape[0]), 4)
    Hz = H.tolil(); Hz[3, 3] = 0; Hz = sp.csr_matrix(Hz); Hz.eliminate_zeros()
    with pytest.raises(RuntimeError, match="-12: ichol: zero on the diagonal"):
        tri_factors(emu, Hz, 4)


@pytest.mark.parametrize("with_T,isolated", [(False, False), (True, False), (False, True), (True, True)])
def test_jk_system(emu, oracle, with_T, isolated):
    """ssn_jk_system's kernels (warp per row, votes): bk1*speye + (T+H0)/tk, pattern exact, values bit for bit."""
    rs = np.random.RandomState(11)
    m, n = 40, 30
    S = rs.random_sample((m, n)) < 0.5                                   # rows longer than a warp: the 32-entry batches and the vote
    if isolated:
        S[5, :] = False; S[:, 9] = False
    p, q = rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5
    H0 = sp.csr_matrix(oracle.ASAt(S.reshape(-1, order="F"), p, q)); H0.sort_indices()
    N = m + n
    bk1, tk = 0.37, 0.81
    t = rs.random_sample(N) if with_T else None
    ptr, idx, val = H0.indptr.astype(np.int32), H0.indices.astype(np.int32), H0.data.astype(np.float64)
    optr = np.zeros(N + 1, np.int32); oidx = np.full(H0.nnz + N, -7, np.int32); oval = np.full(H0.nnz + N, np.nan)
    onnz = C.c_int64(0); err = C.create_string_buffer(256)
    st = emu.emu_jk_system(C.c_int64(m), C.c_int64(n), C.c_int64(H0.nnz), _p(ptr), _p(idx), _p(val), _p(t) if with_T else None,
                           C.c_double(bk1), C.c_double(tk), C.byref(onnz), _p(optr), _p(oidx), _p(oval), err, C.c_int(256))
    assert st == 0, err.value
    Jk = sp.csr_matrix((oval[:onnz.value], oidx[:onnz.value], optr), shape=(N, N))
    TH = H0 if t is None else (sp.diags(t).tocsr() + H0)
    ref = (bk1 * sp.identity(N, format="csr") + TH / tk).tocsr(); ref.sort_indices()
    assert np.array_equal(Jk.indptr, ref.indptr) and np.array_equal(Jk.indices, ref.indices)
    assert np.array_equal(Jk.data, ref.data)
    if isolated:
        assert H0[n + 5, n + 5] == 0 and Jk[n + 5, n + 5] == bk1 + (0.0 if t is None else t[n + 5]) / tk
