"""CPU-only: the C-ABI library loads and exports every symbol include/ssnamg.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "ssnamg.h")).read()
    return sorted(set(re.findall(r"^SSN_API[^;(]*?\b(ssn_\w+)\s*\(", h, flags=re.M)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for name in ["ssn_ax", "ssn_aty", "ssn_asat", "ssn_asatz", "ssn_pcg", "ssn_aug_pcg", "ssn_components",
                 "ssn_hybrid_amg", "ssn_class_amg", "ssn_transfer", "ssn_strength", "ssn_mis_set", "ssn_cf_split",
                 "ssn_mg_vcycle", "ssn_mg_wcycle", "ssn_amg4pot", "ssn_pcg4pot", "ssn_invaat", "ssn_invhht"]:
        assert name in syms


def test_library_exports_every_declared_symbol(ssnamg):
    assert os.path.exists(ssnamg.LIB_PATH), "libssnamg.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(ssnamg.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ssnamg.h but not exported"


def test_python_binding_covers_the_header(ssnamg):
    assert sorted(ssnamg.SIGNATURES) == declared_symbols()
    ssnamg.load()
    lib = ctypes.CDLL(ssnamg.LIB_PATH)
    lib.ssn_version.restype = ctypes.c_int
    assert lib.ssn_version() >= 100


def test_no_cpu_fallback(ssnamg):
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    with pytest.raises(Exception):
        ssnamg.Ax(np.zeros(6), np.ones(2), np.ones(3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "codes-of-ipd-ssn-amg-method_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_mex_shims_type_check_against_the_header():
    """MATLAB is absent, so the MEX shims (one per reference function, mex/*.c) cannot be built here;
    they are at least type-checked against include/ssnamg.h with a declaration-only mex.h stub."""
    import glob
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shims = sorted(glob.glob(os.path.join(root, "mex", "*.c")))
    names = {os.path.splitext(os.path.basename(f))[0] for f in shims}
    for fn in ("Ax", "Aty", "ASAt", "ASAtz", "PCG", "aug_PCG", "components", "Hybrid_AMG", "Class_AMG", "transfer", "strength",
               "mis_set", "cf_split_mex", "MG_Vcycle", "MG_Wcycle", "AMG4POT", "PCG4POT", "invAAt", "invHHt", "warmup_class1",
               "Hybrid_twogrid", "twogrid_bigph", "twogrid", "warmup_class2", "APD_SsN_Class1_mex", "APD_SsN_Class2_mex"):
        assert fn in names, f"no MEX shim for {fn}"
    for f in shims:
        r = subprocess.run([gcc, "-std=c99", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function",
                            "-I" + os.path.join(root, "tests", "mex_stub"), "-I" + os.path.join(root, "mex"), f],
                           capture_output=True, text=True)
        assert r.returncode == 0, f"{os.path.basename(f)}:\n{r.stderr}"


def _build_mex_host(tmp_path):
    """Compiles the functional MEX runtime stand-in, three shims as separate shared objects (two of them the REAL
    mex/MG_Wcycle.c and mex/mis_set.c) and the C host that dlopens them like MATLAB loads .mex files."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    assert gcc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "codes-of-ipd-ssn-amg-method_b200")
    stub = os.path.join(root, "tests", "mex_stub")
    out = str(tmp_path)
    run = lambda cmd: subprocess.run(cmd, capture_output=True, text=True)
    r = run([gcc, "-std=c99", "-D_GNU_SOURCE", "-O1", "-Wall", "-Werror", "-fPIC", "-shared", "-I" + stub,
             os.path.join(stub, "mex_runtime.c"), "-o", os.path.join(out, "libmexrt.so")])
    assert r.returncode == 0, r.stderr
    link = ["-L" + out, "-lmexrt", "-L" + lib, "-lssnamg", "-Wl,-rpath," + out, "-Wl,-rpath," + lib]
    for name, src in (("keep", os.path.join(root, "tests", "c", "shim_keep.c")), ("MG_Wcycle", os.path.join(root, "mex", "MG_Wcycle.c")),
                      ("mis_set", os.path.join(root, "mex", "mis_set.c"))):
        r = run([gcc, "-std=c99", "-O1", "-Wall", "-Werror", "-Wno-unused-function", "-fPIC", "-shared", "-I" + stub,
                 "-I" + os.path.join(root, "mex"), src, "-o", os.path.join(out, name + ".mex.so")] + link)
        assert r.returncode == 0, f"{name}:\n{r.stderr}"
    exe = os.path.join(out, "mex_host")
    r = run([gcc, "-std=c99", "-D_GNU_SOURCE", "-O1", "-Wall", "-Werror", "-I" + stub, os.path.join(root, "tests", "c", "mex_host.c"),
             "-o", exe] + link + ["-ldl", "-lm"])
    assert r.returncode == 0, r.stderr
    return exe, out


def test_mex_shims_link_and_report_a_missing_device(tmp_path, ssnamg):
    """The real shim sources compile and LINK against libssnamg.so (a functional stand-in for MATLAB's MEX runtime
    replaces mex.h's library); without a CUDA device the first shim raises ssnamg:nogpu -- no CPU fallback."""
    import subprocess
    import torch
    exe, out = _build_mex_host(tmp_path)
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present: covered by test_mex_shims_share_one_context")
    r = subprocess.run([exe, out], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stdout, r.stderr)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_mex_shims_share_one_context(tmp_path, gpu):
    """C host (tests/c/mex_host.c), no Python between the shims: three shim shared objects, each with its own static
    handle, share ONE library context -- Class_AMG's hierarchy kept by one shim is cycled by the real MG_Wcycle shim,
    mis_set's random draws are seen through another shim, a clear through one is an error in the other
    (the reference's `global Ack Prok J smoth_it Rk`, AMG/Class_AMG.m:43,110; AMG/MG_Wcycle.m:9; rand, AMG/mis_set.m:35)."""
    import subprocess
    exe, out = _build_mex_host(tmp_path)
    r = subprocess.run([exe, out], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("OK:") and "3 shims, one context" in r.stdout
