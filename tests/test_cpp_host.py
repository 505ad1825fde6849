"""The C++ host layer over the C-ABI (include/tod_b200.hpp): compiles and links with plain g++ against the built
library (CPU), and runs the two cells end to end on a B200 (GPU) from C++, with no Python in the data path."""
import os
import subprocess

import pytest

from conftest import ROOT
from tod_b200 import capi

SRC = os.path.join(ROOT, "tests", "cpp", "cells_smoke.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "cells_smoke")


def build():
    capi.load()
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L", libdir,
           "-ltod_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return EXE


def test_cpp_host_layer_compiles_and_links():
    exe = build()
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libtod_b200.so" in out


@pytest.mark.gpu
def test_cpp_cells_end_to_end():
    exe = EXE if os.path.exists(EXE) else build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK"), r.stdout
