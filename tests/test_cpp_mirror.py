"""The C++ host-side mirror of the reference API (include/tspice_b200.hpp): compiles and links against the
C-ABI library everywhere; runs its rr.cir scenario on a GPU box."""
import os
import subprocess

import pytest

import parity_util as PU

ROOT = PU.ROOT
SRC = os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp")
LIBDIR = os.path.join(ROOT, "toy-spice_b200")


def _build(tmp_path):
    exe = str(tmp_path / "test_mirror")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-ltspice_b200", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_cpp_mirror_compiles_and_links(built, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "--link-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "sm_100a" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_runs_rr(built, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror ok" in r.stdout
