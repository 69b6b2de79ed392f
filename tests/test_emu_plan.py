"""CPU-only: the REAL text of csrc/plan_ops.cu -- Ax, Aty, the fused SsN residual (z, prox, s bit for bit), active-set
compaction, the batched and the screened line-search kernels, the Armijo loop built on them, the fused outer-loop
updates and the fused A-ADMM warm start -- compiled with g++ against tests/emu/common.cuh and compared with the oracle,
as tests/test_gpu_plan.py / tests/test_gpu_driver.py do on the device.  Small plans only (every CUDA thread is a host
thread); shapes chosen to reach the vectorised (m % 16 == 0) and the ragged code paths."""
import ctypes as C

import numpy as np
import pytest

RTOL = 1e-10
from emu_build import slow as _slow                      # SSN_EMU_FULL=1: the cases that cost minutes of host threads
SHAPES = [(1, 1), (7, 5), (64, 48), (45, 130), (258, 9)]     # (258, 9): two full strips -> the cp.async staged batches, and a ragged strip


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    import emu_build
    lib = emu_build.build(tmp_path_factory.mktemp("emu_plan"), "emu_plan.cpp", ["plan_ops.cu", "sparse.cu", "plan_ops.cuh", "sparse.cuh"],
                          "libemu_plan.so")
    lib.emu_error.restype = C.c_char_p
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ok(lib, st):
    assert st == 0, f"{st}: {lib.emu_error().decode()}"


def close(a, b, rtol=RTOL):
    a = np.asarray(a); b = np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return np.all(np.abs(a - b) <= rtol * scale)


def weights(m, n, seed, unit):
    rs = np.random.RandomState(seed)
    return (np.ones(m), np.ones(n)) if unit else (rs.random_sample(m) + 0.5, rs.random_sample(n) + 0.5)


def gama_args(gama, mn):
    """(vector pointer, scalar) as the C ABI takes them"""
    if np.isscalar(gama):
        return None, float(gama)
    return np.ascontiguousarray(gama, dtype=np.float64), float("inf")


def prox_residual(lib, w, lam, p, q, tk, gama, full=True):
    m, n = p.size, q.size
    gv, gs = gama_args(gama, m * n)
    axp = np.zeros(m + n); prox = np.zeros(m * n); z = np.zeros(m * n); s = np.zeros(m * n, np.uint8); scal = np.zeros(2)
    _ok(lib, lib.emu_prox_residual(_p(w), _p(lam), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), _p(gv), C.c_double(gs),
                                   _p(axp) if full else None, _p(prox) if full else None, _p(z) if full else None, _p(s) if full else None, _p(scal)))
    return {"Axprox": axp, "prox": prox, "z": z, "s": s.astype(bool), "norm2": scal[0], "count": int(scal[1])}


@pytest.mark.parametrize("m,n", SHAPES)
@pytest.mark.parametrize("unit", [True, False])
def test_ax_and_aty(emu, oracle, m, n, unit):
    rs = np.random.RandomState(m * 131 + n)
    x = rs.standard_normal(m * n); p, q = weights(m, n, 1, unit)
    y = np.zeros(m + n)
    _ok(emu, emu.emu_ax(_p(x), _p(p), _p(q), C.c_int64(m), C.c_int64(n), _p(y)))
    assert close(y, oracle.Ax(x, p, q))
    yy = rs.standard_normal(n + m); z = np.zeros(m * n)
    _ok(emu, emu.emu_aty(_p(yy), _p(p), _p(q), C.c_int64(m), C.c_int64(n), _p(z)))
    assert np.array_equal(z, oracle.Aty(yy, p, q))                       # bit for bit


@pytest.mark.parametrize("m,n", [(7, 5), (64, 48), (45, 130), (258, 9)])
@pytest.mark.parametrize("gmode", ["inf", "scalar", "vector"])
def test_prox_residual(emu, oracle, m, n, gmode):
    rs = np.random.RandomState(m + 3 * n)
    p, q = weights(m, n, 4, False)
    w = rs.standard_normal(m * n); lam = 0.5 * rs.standard_normal(m + n); tk = 0.37
    gama = {"inf": np.inf, "scalar": 0.8, "vector": rs.random_sample(m * n) + 0.1}[gmode]
    z = 1 / tk * (w - oracle.Aty(lam, p, q))                             # Class1/APD_SsN_Class1.m:139
    s = (z >= 0) & (z <= gama)                                           # :140
    px = np.minimum(np.maximum(0.0, z), gama)                            # :32
    out = prox_residual(emu, w, lam, p, q, tk, gama)
    assert np.array_equal(out["z"], z) and np.array_equal(out["s"], s) and np.array_equal(out["prox"], px)
    assert out["count"] == int(s.sum())
    assert close(out["Axprox"], oracle.Ax(px, p, q))
    term = float(px @ px) if gmode == "inf" else float(z @ z) - float((z - px) @ (z - px))    # APD_SsN_Class1.m:183-187
    assert abs(out["norm2"] - term) <= 1e-12 * float(z @ z) + 1e-300
    lite = prox_residual(emu, w, lam, p, q, tk, gama, full=False)        # line-search form: norm only
    assert abs(lite["norm2"] - out["norm2"]) <= 1e-14 * abs(out["norm2"])


@pytest.mark.parametrize("m,n", [(64, 48), (258, 9)])
def test_prox_residual_unit_weights(emu, oracle, m, n):
    """p = q = 1 (every shipped configuration): the block-uniform unit-weight path of the fused residual gives the bits of the
    general expression `1/tk*(wk - Aty(lk,p,q))`; Class 2 likewise."""
    rs = np.random.RandomState(m + 7 * n)
    p, q = np.ones(m), np.ones(n)
    w = rs.standard_normal(m * n); lam = 0.5 * rs.standard_normal(m + n); tk = 0.37
    z = 1 / tk * (w - oracle.Aty(lam, p, q))
    out = prox_residual(emu, w, lam, p, q, tk, np.inf)
    assert np.array_equal(out["z"], z) and np.array_equal(out["s"], z >= 0) and np.array_equal(out["prox"], np.maximum(z, 0.0))
    assert close(out["Axprox"], oracle.Ax(np.maximum(z, 0.0), p, q))
    N = m + n; mn = m * n
    w2 = rs.standard_normal(mn + N); lam2 = 0.5 * rs.standard_normal(N + 1); phi = rs.random_sample(mn) + 0.5
    z2 = 1 / tk * (w2 - np.concatenate([oracle.Aty(lam2[:N], p, q) + lam2[N] * phi, lam2[:N]]))
    hp = np.zeros(N + 1); prox = np.zeros(mn + N); s = np.zeros(mn, np.uint8); t = np.zeros(N); scal = np.zeros(3)
    _ok(emu, emu.emu_prox_residual_pot(_p(w2), _p(lam2), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), _p(phi),
                                       _p(hp), _p(prox), _p(s), _p(t), _p(scal)))
    assert np.array_equal(s.astype(bool), z2[:mn] >= 0) and np.array_equal(prox, np.maximum(z2, 0.0))


@pytest.mark.parametrize("m,n", [(7, 5), (64, 48), (45, 130)])
@pytest.mark.parametrize("unit_phi", [True, False])
def test_prox_residual_pot(emu, oracle, m, n, unit_phi):
    """The fused residual of partial OT (plan_reduce_kernel<PROX, G_PHI> + the slack kernel): z-derived flags and prox bit
    for bit against Class2/APD_SsN_Class2.m:124-130, H*prox to rounding."""
    rs = np.random.RandomState(5 * m + n)
    p, q = weights(m, n, 6, False)
    N = m + n; mn = m * n
    w = rs.standard_normal(mn + N); lam = 0.5 * rs.standard_normal(N + 1); tk = 0.41
    phi = np.ones(mn) if unit_phi else rs.random_sample(mn) + 0.5
    z = 1 / tk * (w - np.concatenate([oracle.Aty(lam[:N], p, q) + lam[N] * phi, lam[:N]]))
    pz = np.maximum(z, 0.0)
    Hp = np.concatenate([oracle.Ax(pz[:mn], p, q) + pz[mn:], [phi @ pz[:mn]]])
    hp = np.zeros(N + 1); prox = np.zeros(mn + N); s = np.zeros(mn, np.uint8); t = np.zeros(N); scal = np.zeros(3)
    _ok(emu, emu.emu_prox_residual_pot(_p(w), _p(lam), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), _p(phi),
                                       _p(hp), _p(prox), _p(s), _p(t), _p(scal)))
    assert np.array_equal(s.astype(bool), z[:mn] >= 0) and np.array_equal(t > 0.5, z[mn:] >= 0)
    assert np.array_equal(prox, pz)
    assert int(scal[1]) == int((z[:mn] >= 0).sum())
    assert close(hp, Hp)
    assert abs(scal[0] - float(pz @ pz)) <= 1e-12 * float(pz @ pz)
    scal2 = np.zeros(3)
    _ok(emu, emu.emu_prox_residual_pot(_p(w), _p(lam), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), _p(phi),
                                       None, None, None, None, _p(scal2)))
    assert abs(scal2[0] - scal[0]) <= 1e-14 * abs(scal[0])


@pytest.mark.parametrize("m,n,density", [(5, 4, 0.5), (64, 48, 0.1), (45, 130, 0.05), (48, 7, 1.0), (33, 20, 0.0)])
def test_active_set_compaction(emu, m, n, density):
    """Y = sparse(reshape(s,m,n)) as CSC coordinate lists, rows ascending inside a column (ASAt.m:15)"""
    rs = np.random.RandomState(m * n)
    S = rs.random_sample((m, n)) < density
    s = np.ascontiguousarray(S.reshape(-1, order="F"), dtype=np.uint8)
    colptr = np.zeros(n + 1, np.int32); yrow = np.zeros(m * n + 1, np.int32); ycol = np.zeros(m * n + 1, np.int32); rowcount = np.zeros(m, np.int32)
    E = C.c_int64(0)
    _ok(emu, emu.emu_active_set(_p(s), C.c_int64(m), C.c_int64(n), _p(colptr), _p(yrow), _p(ycol), _p(rowcount), C.byref(E)))
    jj, ii = np.nonzero(S.T)                                             # column-major order == find(s)
    assert E.value == ii.size
    assert np.array_equal(yrow[:E.value], ii) and np.array_equal(ycol[:E.value], jj)
    assert np.array_equal(colptr, np.concatenate([[0], np.cumsum(S.sum(axis=0))]))
    assert np.array_equal(rowcount, S.sum(axis=1))


def _ls_state(oracle, m, n, seed, unit=True):
    rs = np.random.RandomState(seed)
    p, q = weights(m, n, seed, unit)
    w = rs.standard_normal(m * n) - 0.3; lam = 0.3 * rs.standard_normal(m + n); zeta = rs.standard_normal(m + n)
    wlk = rs.standard_normal(m + n)
    return p, q, w, lam, zeta, wlk


@pytest.mark.parametrize("m,n,unit", [(64, 48, True), pytest.param(45, 70, False, marks=_slow)])
def test_trial_kernels_agree_with_single_evaluations(emu, oracle, m, n, unit):
    """plan_prox_trials (8 trial vectors per read of w) and plan_prox_trials_lin (screened, lam + delta^ll*zeta) against
    the norm of the fused residual evaluated trial by trial."""
    p, q, w, lam, zeta, wlk = _ls_state(oracle, m, n, 5, unit)
    tk, delta = 0.6, 0.9
    for nt in (3, 8):
        lamT = np.ascontiguousarray(np.stack([lam + delta ** t * zeta for t in range(nt)]))
        ref = np.array([prox_residual(emu, w, lamT[t], p, q, tk, np.inf, full=False)["norm2"] for t in range(nt)])
        px = [np.maximum((w - oracle.Aty(lamT[t], p, q)) / tk, 0.0) for t in range(nt)]
        assert np.allclose(ref, [v @ v for v in px], rtol=1e-12)
        out = np.zeros(nt)
        _ok(emu, emu.emu_prox_trials(_p(w), _p(lamT), C.c_int(nt), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), None,
                                     C.c_double(float("inf")), _p(out)))
        assert np.allclose(out, ref, rtol=1e-13, atol=0)
    for ll0, nt in ((0, 1), (2, 16)):
        out = np.zeros(nt + 1)
        _ok(emu, emu.emu_prox_trials_lin(_p(w), _p(lam), _p(zeta), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), C.c_double(delta),
                                         C.c_int(ll0), C.c_int(nt), _p(out)))
        ref = [prox_residual(emu, w, lam + delta ** (ll0 + t) * zeta, p, q, tk, np.inf, full=False)["norm2"] for t in range(nt)]
        assert np.allclose(out[:nt], ref, rtol=1e-12, atol=1e-300)
        assert 0 <= out[nt] <= m * n                                       # entries that survived the screen


@pytest.mark.parametrize("m,n,gama,batch", [pytest.param(64, 48, np.inf, 0, marks=_slow), pytest.param(45, 70, 2.5, 0, marks=_slow),
                                            pytest.param(45, 70, np.inf, 4, marks=_slow)])
def test_linesearch_matches_the_reference_loop(emu, oracle, m, n, gama, batch):
    """Class1/APD_SsN_Class1.m:182-211 trial by trial (oracle operators) against the device loop (adaptive / batched) along
    an over-long steepest-descent direction, which needs a handful of backtracking steps."""
    p, q, w, lam, _, wlk = _ls_state(oracle, m, n, 9)
    tk, bk1, nu, delta, ll_max = 0.6, 0.25, 0.2, 0.9, 500
    zof = lambda l: (w - oracle.Aty(l, p, q)) / tk
    prox = lambda l: np.minimum(np.maximum(zof(l), 0.0), gama)
    # APD_SsN_Class1.m:183-187: ||prox(z)||^2 for gama = Inf (prob < 3), ||z||^2 - ||z - prox(z)||^2 with capacities (prob = 3)
    term = (lambda l: np.sum(prox(l) ** 2)) if np.isinf(gama) else (lambda l: np.sum(zof(l) ** 2) - np.sum((zof(l) - prox(l)) ** 2))
    cF = lambda l: bk1 / 2 * (l @ l) - wlk @ l + 0.5 * tk * term(l)
    cF_old = cF(lam)
    Fk = bk1 * lam - oracle.Ax(prox(lam), p, q) - wlk                      # the gradient of cF at lam (where no entry sits at its capacity)
    zeta = -0.1 * Fk
    ress = abs(float(Fk @ zeta))
    ll_ref = 0
    while cF(lam + delta ** ll_ref * zeta) > cF_old - nu * delta ** ll_ref * ress and ll_ref < ll_max:
        ll_ref += 1
    assert 3 <= ll_ref <= 60, ll_ref
    gv, gs = gama_args(gama, m * n)
    lam_new = np.zeros(m + n); ll = C.c_int(-1); n2 = C.c_double(0); cFo = C.c_double(0); passes = C.c_int(0)
    _ok(emu, emu.emu_linesearch(_p(w), _p(lam), _p(zeta), _p(wlk), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), C.c_double(bk1),
                                _p(gv), C.c_double(gs), C.c_double(nu), C.c_double(delta), C.c_int(ll_max), C.c_double(cF_old), C.c_double(ress),
                                C.c_int(batch), _p(lam_new), C.byref(ll), C.byref(n2), C.byref(cFo), C.byref(passes)))
    assert ll.value == ll_ref
    assert np.array_equal(lam_new, lam + delta ** ll_ref * zeta)
    assert abs(cFo.value - cF(lam_new)) <= 1e-11 * abs(cF(lam_new))
    assert 1 <= passes.value <= ll_ref + 1


@pytest.mark.parametrize("m,n,gama", [(64, 48, np.inf), (45, 70, 0.3)])
def test_fused_outer_loop_updates(emu, oracle, m, n, gama):
    """ssn_apd_begin / ssn_apd_end against the script lines (Class1/APD_SsN_Class1.m:125-126, 239-254)"""
    rs = np.random.RandomState(3)
    p, q = weights(m, n, 3, False)
    c = rs.random_sample(m * n); xk = rs.random_sample(m * n); vk = rs.random_sample(m * n); lam = rs.standard_normal(m + n)
    ak, bk, tk = 1.3, 0.8, 0.55
    wk = np.zeros(m * n); axk = np.zeros(m + n)
    _ok(emu, emu.emu_apd_begin(_p(c), _p(xk), _p(vk), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(ak), C.c_double(bk), _p(wk), _p(axk)))
    wk_ref = -c + bk * (xk + ak * vk) / ak ** 2
    assert np.allclose(wk, wk_ref, rtol=1e-14, atol=1e-15) and close(axk, oracle.Ax(xk, p, q))
    gv, gs = gama_args(gama, m * n)
    xk1 = np.zeros(m * n); vk1 = np.zeros(m * n); axk1 = np.zeros(m + n); scal = np.zeros(2)
    _ok(emu, emu.emu_apd_end(_p(c), _p(wk), _p(xk), _p(lam), _p(p), _p(q), C.c_int64(m), C.c_int64(n), C.c_double(tk), C.c_double(ak), _p(gv),
                             C.c_double(gs), _p(xk1), _p(vk1), _p(axk1), _p(scal)))
    prox = lambda v: np.minimum(np.maximum(v, 0.0), gama)
    x1 = prox((wk - oracle.Aty(lam, p, q)) / tk)
    assert np.allclose(xk1, x1, rtol=1e-14, atol=1e-15)
    assert np.allclose(vk1, x1 + (x1 - xk) / ak, rtol=1e-13, atol=1e-14)
    assert close(axk1, oracle.Ax(x1, p, q))
    assert abs(scal[0] - c @ x1) <= 1e-11 * abs(c @ x1)
    kx = x1 - prox(x1 - c - oracle.Aty(lam, p, q))
    assert abs(scal[1] - kx @ kx) <= 1e-10 * max(kx @ kx, 1e-30)


@pytest.mark.parametrize("m,n,gama,unit", [pytest.param(48, 40, np.inf, True, marks=_slow), (45, 70, 0.05, False)])
def test_fused_warm_start(emu, oracle, m, n, gama, unit):
    """ssn_warmup_class1 (two fused plan-wide kernels per A-ADMM iteration) against the oracle's line-by-line restatement
    of Class1/warmup_class1.m"""
    from oracle import driver as odrv
    rs = np.random.RandomState(5)
    p, q = weights(m, n, 6, unit)
    c = rs.random_sample(m * n); l = rs.random_sample(m) + 0.1; r = rs.random_sample(n) + 0.1
    r = r * (l.sum() / r.sum())
    its = 12
    x_ref, l_ref = odrv.warmup_class1(c, r, l, p, q, gama, 0, its)
    b = np.concatenate([r, l])
    gv, gs = gama_args(gama, m * n)
    xk = np.zeros(m * n); lk = np.zeros(m + n)
    _ok(emu, emu.emu_warmup_class1(_p(c), _p(b), _p(p), _p(q), C.c_int64(m), C.c_int64(n), _p(gv), C.c_double(gs), C.c_int(its), _p(xk), _p(lk)))
    assert np.linalg.norm(xk - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    assert np.linalg.norm(lk - l_ref) <= 1e-9 * np.linalg.norm(l_ref)


@pytest.mark.parametrize("m,n,unit_phi", [(64, 48, True), (45, 70, False)])
def test_fused_outer_loop_updates_partial_ot(emu, oracle, m, n, unit_phi):
    """ssn_apd_begin_pot / ssn_apd_end_pot against the script lines (Class2/APD_SsN_Class2.m:121-122, 231-238)"""
    rs = np.random.RandomState(7)
    p, q = weights(m, n, 3, False)
    N = m + n; mn = m * n
    c = rs.random_sample(mn); uk = rs.random_sample(mn + N); vk = rs.random_sample(mn + N); lk = 0.3 * rs.standard_normal(N + 1)
    phi = np.ones(mn) if unit_phi else rs.random_sample(mn) + 0.5
    b = np.concatenate([rs.random_sample(N) + 0.1, [0.7]])
    ak, bk, tk = 1.3, 0.8, 0.55; bk1 = bk / (1 + ak)
    H = lambda u: np.concatenate([oracle.Ax(u[:mn], p, q) + u[mn:], [phi @ u[:mn]]])
    Ht = lambda lam: np.concatenate([oracle.Aty(lam[:N], p, q) + lam[N] * phi, lam[:N]])
    wc = np.concatenate([c, np.zeros(N)])
    wk = np.zeros(mn + N); huk = np.zeros(N + 1); wlk = np.zeros(N + 1)
    _ok(emu, emu.emu_apd_begin_pot(_p(c), _p(uk), _p(vk), _p(p), _p(q), C.c_int64(m), C.c_int64(n), _p(phi), _p(b), _p(lk), C.c_double(ak),
                                   C.c_double(bk), C.c_double(bk1), _p(wk), _p(huk), _p(wlk)))
    wk_ref = -wc + bk * (uk + ak * vk) / ak ** 2
    assert np.allclose(wk, wk_ref, rtol=1e-14, atol=1e-15)
    assert close(huk, H(uk))
    assert close(wlk, bk1 * (lk - 1 / bk * (H(uk) - b)) - b)
    uk1 = np.zeros(mn + N); vk1 = np.zeros(mn + N); huk1 = np.zeros(N + 1); scal = np.zeros(5)
    _ok(emu, emu.emu_apd_end_pot(_p(c), _p(wk), _p(uk), _p(lk), _p(p), _p(q), C.c_int64(m), C.c_int64(n), _p(phi), _p(b), C.c_double(tk),
                                 C.c_double(ak), _p(uk1), _p(vk1), _p(huk1), _p(scal)))
    u1 = np.maximum((wk - Ht(lk)) / tk, 0.0)
    assert np.allclose(uk1, u1, rtol=1e-14, atol=1e-15)
    assert np.allclose(vk1, u1 + (u1 - uk) / ak, rtol=1e-13, atol=1e-14)
    assert close(huk1, H(u1))
    x1, y1, z1 = u1[:mn], u1[mn:mn + n], u1[mn + n:]
    ref = [c @ x1,
           np.sum((x1 - np.maximum(x1 - c - Ht(lk)[:mn], 0.0)) ** 2),
           np.sum((y1 - np.maximum(y1 - lk[:n], 0.0)) ** 2),
           np.sum((z1 - np.maximum(z1 - lk[n:N], 0.0)) ** 2),
           np.sum((H(u1) - b) ** 2)]
    for got, want in zip(scal, ref):
        assert abs(got - want) <= 1e-10 * max(abs(want), 1e-30)


@pytest.mark.parametrize("m,n,unit", [pytest.param(48, 40, True, marks=_slow), (45, 70, False)])
def test_fused_warm_start_partial_ot(emu, oracle, m, n, unit):
    """ssn_warmup_class2 (two fused plan-wide kernels + two one-block kernels per A-ADMM iteration) against the oracle's
    line-by-line restatement of Class2/warmup_class2.m"""
    from oracle import driver as odrv
    rs = np.random.RandomState(5)
    p, q = weights(m, n, 6, unit)
    mn = m * n; N = m + n
    c = rs.random_sample(mn); l = rs.random_sample(m) + 0.1; r = rs.random_sample(n) + 0.1
    phi = np.ones(mn) if unit else rs.random_sample(mn) + 0.5
    mu = 0.6 * min(l.sum(), r.sum())
    its = 12
    u_ref, l_ref = odrv.warmup_class2(c, r, l, p, q, mu, phi, 0, its)
    b = np.concatenate([r, l, [mu]])
    uk = np.zeros(mn + N); lk = np.zeros(N + 1)
    _ok(emu, emu.emu_warmup_class2(_p(c), _p(b), _p(p), _p(q), C.c_int64(m), C.c_int64(n), _p(phi), C.c_int(its), _p(uk), _p(lk)))
    assert np.linalg.norm(uk - np.asarray(u_ref).ravel()) <= 1e-9 * np.linalg.norm(u_ref)
    assert np.linalg.norm(lk - np.asarray(l_ref).ravel()) <= 1e-9 * np.linalg.norm(l_ref)
