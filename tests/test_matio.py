"""CPU-only: the C-side MAT-file level 5 reader (include/ssnamg_io.h, csrc/matv5.c) against SciPy's
reader, against hand-built files that use MATLAB's storage tricks (integer-compressed doubles, small
data elements, zlib-compressed variables), and -- in the build container, where the reference is
present -- against the two bundled problem files."""
import ctypes
import importlib
import os
import re
import struct
import zlib

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from conftest import ROOT

BUNDLED = ["/root/reference/Class1/InputData/data1-500.mat", "/root/reference/Class2/InputData/data4-500.mat"]


@pytest.fixture(scope="module")
def mio():
    return importlib.import_module("codes-of-ipd-ssn-amg-method_b200.matio")


def test_library_exports_every_declared_symbol(mio):
    h = open(os.path.join(ROOT, "include", "ssnamg_io.h")).read()
    declared = sorted(set(re.findall(r"^SSN_IO_API[^;(]*?\b(ssn_\w+)\s*\(", h, flags=re.M)))
    assert len(declared) == 10 and sorted(mio.SIGNATURES) == declared
    assert os.path.exists(mio.LIB_PATH), "libssnmat.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(mio.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


# ---------------------------------------------------------------- a minimal MAT-5 writer for the tests
MI = {"int8": 1, "uint8": 2, "int16": 3, "uint16": 4, "int32": 5, "uint32": 6, "float32": 7, "float64": 9,
      "int64": 12, "uint64": 13}
MX_DOUBLE = 6


def _elem(mi_type, payload, small_ok=True):
    if small_ok and 0 < len(payload) <= 4:                       # small data element: size in the high half-word
        return struct.pack("<HH", mi_type, len(payload)) + payload + b"\0" * (4 - len(payload))
    pad = (-len(payload)) % 8
    return struct.pack("<II", mi_type, len(payload)) + payload + b"\0" * pad


def _matrix(name, arr, stored, mx_class=MX_DOUBLE, compress=True, flags=0):
    arr = np.asarray(arr)
    if arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    body = _elem(6, struct.pack("<II", mx_class | flags, 0), small_ok=False)
    body += _elem(5, struct.pack("<ii", *arr.shape), small_ok=False)
    body += _elem(1, name.encode())
    body += _elem(MI[stored], np.asfortranarray(arr).astype(stored).tobytes(order="F"))
    el = struct.pack("<II", 14, len(body)) + body
    if compress:
        z = zlib.compress(el)
        return struct.pack("<II", 15, len(z)) + z              # compressed elements are not padded
    return el


def _file(elements):
    head = b"MATLAB 5.0 MAT-file, Platform: MACI64, written by tests/test_matio.py".ljust(116) + b"\0" * 8
    return head + struct.pack("<H", 0x0100) + b"IM" + b"".join(elements)


def test_matlab_storage_tricks(mio, tmp_path):
    """double-class variables whose data MATLAB stored as uint8 / uint16 / int32 / single, scalars in the small
    data element form, compressed and uncompressed elements mixed in one file."""
    rs = np.random.RandomState(0)
    ones = np.ones(37)
    ints = rs.randint(-30000, 30000, size=(5, 7)).astype(np.float64)
    big = rs.randint(-2 ** 31, 2 ** 31 - 1, size=11).astype(np.float64)
    sing = rs.random_sample((3, 4)).astype(np.float32).astype(np.float64)
    dbl = rs.standard_normal((6, 2))
    dbl[0, 0] = np.inf
    path = tmp_path / "tricks.mat"
    path.write_bytes(_file([
        _matrix("p", ones, "uint8"), _matrix("m", [500.0], "uint16"), _matrix("tiny", [3.0], "uint8"),
        _matrix("ints", ints, "int16", compress=False), _matrix("big", big, "int32"),
        _matrix("sing", sing, "float32"), _matrix("dbl", dbl, "float64", compress=False),
        _matrix("a_long_variable_name_of_31_chars", [1.0, 2.0], "uint8"),
    ]))
    with mio.MatFile(str(path)) as mf:
        names = [w[0] for w in mf.whos()]
        assert names == ["p", "m", "tiny", "ints", "big", "sing", "dbl", "a_long_variable_name_of_31_chars"]
        assert all(w[2] for w in mf.whos())
        assert np.array_equal(mf.read("p"), ones.reshape(-1, 1))
        assert mf.read("m")[0, 0] == 500.0 and mf.read("tiny")[0, 0] == 3.0
        assert np.array_equal(mf.read("ints"), ints)
        assert np.array_equal(mf.read("big").ravel(), big)
        assert np.array_equal(mf.read("sing"), sing)
        assert np.array_equal(mf.read("dbl"), dbl)
        assert "p" in mf and "nope" not in mf
        with pytest.raises(KeyError):
            mf.read("nope")
    # SciPy reads the same bytes the same way
    d = scipy.io.loadmat(str(path), mat_dtype=True)
    assert np.array_equal(d["ints"], ints) and np.array_equal(d["p"].ravel(), ones)


@pytest.mark.parametrize("compress", [False, True])
def test_against_scipy_writer(mio, tmp_path, compress):
    rs = np.random.RandomState(1)
    vars_ = {"c": rs.random_sample((12, 1)), "r": rs.random_sample(4) + 0.1, "u8": np.arange(200, dtype=np.uint8).reshape(20, 10),
             "i64": np.array([[-2 ** 40, 7]], dtype=np.int64), "f32": rs.random_sample((2, 3)).astype(np.float32),
             "empty": np.zeros((0, 3)), "scalar": 2.5}
    path = tmp_path / "scipy.mat"
    scipy.io.savemat(str(path), vars_, do_compression=compress)
    d = scipy.io.loadmat(str(path))
    with mio.MatFile(str(path)) as mf:
        for name, shape, ok in mf.whos():
            assert ok and shape == d[name].shape, name
            assert np.array_equal(mf.read(name), d[name].astype(np.float64)), name


def test_unsupported_variables_are_reported_not_misread(mio, tmp_path):
    path = tmp_path / "mixed.mat"
    scipy.io.savemat(str(path), {"S": sp.identity(4, format="csc"), "txt": "hello", "st": {"a": 1.0}, "cplx": np.array([1 + 2j]),
                                 "nd": np.zeros((2, 3, 4)), "ok": np.arange(6.0).reshape(2, 3)})
    with mio.MatFile(str(path)) as mf:
        w = {name: ok for name, _, ok in mf.whos()}
        assert w == {"S": False, "txt": False, "st": False, "cplx": False, "nd": False, "ok": True}
        for name in ("S", "txt", "st", "cplx", "nd"):
            with pytest.raises(mio.MatError) as e:
                mf.read(name)
            assert e.value.status == "SSN_MAT_E_UNSUPPORTED"
        assert np.array_equal(mf.read("ok"), np.arange(6.0).reshape(2, 3))


def test_error_statuses(mio, tmp_path):
    with pytest.raises(mio.MatError) as e:
        mio.MatFile(str(tmp_path / "missing.mat"))
    assert e.value.status == "SSN_MAT_E_IO"
    bad = tmp_path / "bad.mat"
    bad.write_bytes(b"not a mat file" * 20)
    with pytest.raises(mio.MatError) as e:
        mio.MatFile(str(bad))
    assert e.value.status == "SSN_MAT_E_FORMAT"
    v73 = tmp_path / "v73.mat"                                      # HDF5-based files start with a different text
    v73.write_bytes(b"MATLAB 7.3 MAT-file, Platform: GLNXA64".ljust(128))
    with pytest.raises(mio.MatError) as e:
        mio.MatFile(str(v73))
    assert e.value.status == "SSN_MAT_E_FORMAT"
    good = _file([_matrix("x", np.arange(100.0), "float64", compress=False)])
    trunc = tmp_path / "trunc.mat"
    trunc.write_bytes(good[:-40])
    with pytest.raises(mio.MatError) as e:
        mio.MatFile(str(trunc))
    assert e.value.status == "SSN_MAT_E_FORMAT"
    z = bytearray(_file([_matrix("x", np.arange(100.0), "float64", compress=True)]))
    z[128 + 8 + 20] ^= 0xFF                                          # corrupt the deflate stream
    corrupt = tmp_path / "corrupt.mat"
    corrupt.write_bytes(bytes(z))
    with pytest.raises(mio.MatError) as e:
        mio.MatFile(str(corrupt))
    assert e.value.status in ("SSN_MAT_E_ZLIB", "SSN_MAT_E_FORMAT")
    big_endian = bytearray(good)
    big_endian[126:128] = b"MI"
    be = tmp_path / "be.mat"
    be.write_bytes(bytes(big_endian))
    with pytest.raises(mio.MatError) as e:
        mio.MatFile(str(be))
    assert e.value.status == "SSN_MAT_E_FORMAT"


def _write_problem(path, m, n, class2=False, with_mn=True, bad=None):
    rs = np.random.RandomState(5)
    P = {"c": rs.random_sample(m * n), "l": rs.random_sample(m) + 0.1, "r": rs.random_sample(n) + 0.1,
         "p": np.ones(m), "q": np.ones(n)}
    els = [_matrix("c", P["c"], "float64"), _matrix("l", P["l"], "float64"), _matrix("r", P["r"], "float64"),
           _matrix("p", P["p"], "uint8"), _matrix("q", P["q"], "uint8")]
    if with_mn:
        els += [_matrix("m", [float(m)], "uint16"), _matrix("n", [float(n)], "uint16")]
    if class2:
        P["phi"], P["mu"] = np.ones(m * n), 0.65 * min(P["l"].sum(), P["r"].sum())
        els += [_matrix("phi", P["phi"], "uint8"), _matrix("mu", [P["mu"]], "float64")]
    else:
        els += [_matrix("gama", np.full(m * n, np.inf), "float64")]
    if bad == "short_c":
        els[0] = _matrix("c", P["c"][:-1], "float64")
    if bad == "no_r":
        del els[2]
    path.write_bytes(_file(els))
    return P


@pytest.mark.parametrize("class2", [False, True])
@pytest.mark.parametrize("with_mn", [True, False])
def test_problem_load(mio, tmp_path, class2, with_mn):
    path = tmp_path / "prob.mat"
    P = _write_problem(path, 7, 5, class2=class2, with_mn=with_mn)
    Q = mio.load_problem(str(path))
    assert (Q["m"], Q["n"]) == (7, 5) and Q["gama"] == np.inf
    for k in ("c", "l", "r", "p", "q"):
        assert np.array_equal(Q[k], P[k]), k
    if class2:
        assert np.array_equal(Q["phi"], P["phi"]) and Q["mu"] == P["mu"]
    else:
        assert "phi" not in Q and "mu" not in Q


@pytest.mark.parametrize("bad", ["short_c", "no_r"])
def test_problem_load_rejects_inconsistent_files(mio, tmp_path, bad):
    path = tmp_path / "prob.mat"
    _write_problem(path, 7, 5, bad=bad)
    with pytest.raises(mio.MatError) as e:
        mio.load_problem(str(path))
    assert e.value.status == "SSN_MAT_E_INVALID"


@pytest.mark.parametrize("path", BUNDLED)
def test_bundled_files(mio, path):
    """The two input files the reference ships (absent on the GPU box: skipped there)."""
    if not os.path.exists(path):
        pytest.skip("reference not present")
    d = scipy.io.loadmat(path, mat_dtype=True)
    with mio.MatFile(path) as mf:
        names = [w[0] for w in mf.whos()]
        assert names == [k for k in d if not k.startswith("__")]
        for name in names:
            a = mf.read(name)
            assert a.shape == d[name].shape and np.array_equal(a, d[name]), name
    P = mio.load_problem(path)
    assert P["m"] == P["n"] == 500 and P["c"].shape == (250000,)
    assert abs(P["r"].sum() - P["l"].sum()) < 1e-9 or "mu" in P
    assert np.all(P["p"] == 1) and np.all(P["q"] == 1)


def test_problems_loader_uses_the_c_reader(tmp_path):
    """problems.load_bundled_class1 goes through libssnmat.so (no SciPy on the product path)."""
    import ssnamg
    path = tmp_path / "prob.mat"
    P = _write_problem(path, 6, 4)
    Q = ssnamg.problems.load_bundled_class1(str(path))
    assert Q["m"] == 6 and Q["n"] == 4 and Q["gama"] == np.inf
    assert np.array_equal(Q["c"], P["c"]) and np.array_equal(Q["r"], P["r"])
    with pytest.raises(FileNotFoundError):
        ssnamg.problems.load_bundled_class1(str(tmp_path / "nope.mat"))


def _build_example(tmp_path):
    import shutil
    import subprocess
    pkg = os.path.join(ROOT, "codes-of-ipd-ssn-amg-method_b200")
    exe = str(tmp_path / "warmstart_from_mat")
    r = subprocess.run([shutil.which("gcc"), "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "warmstart_from_mat.c"), "-o", exe, "-L" + pkg, "-lssnamg", "-lssnmat", "-lm",
                        "-Wl,-rpath," + pkg], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_standalone_c_program_links_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/warmstart_from_mat.c (no MATLAB, no Python): compiles against both headers, reads the problem file,
    and -- without a CUDA device -- stops at ssn_create with a non-zero exit status (no CPU fallback)."""
    import subprocess
    import torch
    exe = _build_example(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "missing.mat")], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot read the file" in r.stderr
    path = tmp_path / "prob.mat"
    _write_problem(path, 9, 6)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_standalone_c_program_on_the_gpu")
    r = subprocess.run([exe, str(path), "5"], capture_output=True, text=True)
    assert r.returncode == 3 and "ssn_create" in r.stderr and "m = 9, n = 6" in r.stdout


@pytest.mark.gpu
def test_standalone_c_program_on_the_gpu(tmp_path, gpu):
    """The same program on the device against the oracle's warm start (Class1/warmup_class1.m) of the same file."""
    import subprocess
    import oracle
    from oracle import driver as odrv
    exe = _build_example(tmp_path)
    path = tmp_path / "prob.mat"
    P = _write_problem(path, 40, 30)
    P["r"] = P["r"] * (P["l"].sum() / P["r"].sum())
    els = [_matrix(k, P[k], "float64") for k in ("c", "l", "r", "p", "q")] + [_matrix("gama", np.full(1200, np.inf), "float64")]
    path.write_bytes(_file(els))
    r = subprocess.run([exe, str(path), "20"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mt = re.search(r"= ([0-9.e+-]+), c'xk = ([0-9.e+-]+), kernels launched: (\d+)", r.stdout)
    assert mt, r.stdout
    xk, _ = odrv.warmup_class1(P["c"], P["r"], P["l"], P["p"], P["q"], np.inf, 0, 20)
    b = np.concatenate([P["r"], P["l"]])
    res = np.linalg.norm(oracle.Ax(xk, P["p"], P["q"]) - b) / np.linalg.norm(b)
    assert abs(float(mt.group(1)) - res) <= 2e-3 * res + 1e-12           # printed with 4 digits
    assert abs(float(mt.group(2)) - P["c"] @ xk) <= 1e-6 * abs(P["c"] @ xk)
    assert int(mt.group(3)) > 0
