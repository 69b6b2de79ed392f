"""Builds a host emulation library: the real csrc/ sources copied next to tests/emu/common.cuh (the stand-in for the
CUDA names) and compiled with g++.  Test infrastructure only."""
import ctypes
import os

import pytest
import re
import shutil
import subprocess

from conftest import ROOT

FULL = os.environ.get("SSN_EMU_FULL") == "1"          # the larger cases cost minutes of host-thread emulation: on request
slow = pytest.mark.skipif(not FULL, reason="set SSN_EMU_FULL=1 (minutes of host-thread emulation)")

CSRC = os.path.join(ROOT, "codes-of-ipd-ssn-amg-method_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


def build(tmpdir, harness, sources, so_name):
    """Copies tests/emu/* and the named real sources into ``tmpdir``, compiles ``harness`` (+ every ``.cu`` in
    ``sources`` that the harness does not #include itself is compiled as its own translation unit) and loads it."""
    d = str(tmpdir)
    shutil.copytree(EMU, d, dirs_exist_ok=True)
    units = [harness]
    for f in sources:
        txt = open(os.path.join(CSRC, f)).read()
        # the one mechanical edit: dynamic shared memory becomes a pointer to the emulator's buffer
        txt = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::dyn_smem;", txt)
        open(os.path.join(d, f), "w").write(txt)
    harness_txt = open(os.path.join(d, harness)).read()
    for f in sources:
        if f.endswith(".cu") and f'#include "{f}"' not in harness_txt:
            units.append(f)
    so = os.path.join(d, so_name)
    cmd = ["g++", "-std=c++20", "-O0", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w", "-I" + d,
           "-I" + os.path.join(ROOT, "include")]
    if os.environ.get("SSN_EMU_ASAN") == "1":              # out-of-bounds accesses of the emulated kernels (run with LD_PRELOAD=libasan.so)
        cmd += ["-fsanitize=address", "-fno-omit-frame-pointer", "-g"]
    for u in units:
        cmd += ["-x", "c++", u]
    r = subprocess.run(cmd + ["-o", so], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]
    lib = ctypes.CDLL(so)
    if not lib.emu_probe_threads(ctypes.c_int(1024)):                # one host thread per CUDA thread of a block
        pytest.skip("this machine cannot create 1024 host threads (ulimit -u?): the kernel emulation needs them")
    return lib
