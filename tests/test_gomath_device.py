"""The device restatements of Go's math.Sin / Max / Min / Pow (csrc/device/models.cuh) against the oracle's host
restatements, bit for bit, on a GPU (compiled with --fmad=false like the strict kernels)."""
import math
import os
import subprocess

import numpy as np
import pytest

import parity_util as PU

ROOT = PU.ROOT
SRC = os.path.join(ROOT, "tests", "cuda", "gomath_device.cu")


def _build(tmp_path):
    exe = str(tmp_path / "gomath_device")
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "--fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-diag-suppress=550", "-o", exe, SRC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_device_gomath_program_compiles_for_sm100a(tmp_path):
    _build(tmp_path)


def test_fractional_pow_follows_go_on_the_host(tmp_path):
    """math.Pow with a fractional exponent other than +-0.5 (MOSFET Level 2 / 3 mobility degradation, junction capacitances
    with MJ != 0.5): Go computes Exp(yf * Log(x)) times the integer power (pow.go), not the correctly rounded power of
    libm.  tsb_go_pow, compiled for the host as the symbolic pass and tests/host_emul.py compile it, reproduces the oracle's
    restatement bit for bit; on the device the same operations run on CUDA's exp / log (each within 1 ulp of glibc's)."""
    src = tmp_path / "p.cpp"
    src.write_text('''#include <cstdio>
#include <cmath>
#include "%s"
int main() { double x, y; while (scanf("%%la %%la", &x, &y) == 2) printf("%%a\\n", tsb_go_pow(x, y)); return 0; }
''' % os.path.join(ROOT, "toy-spice_b200", "csrc", "device", "models.cuh"))
    exe = str(tmp_path / "p")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-o", exe, str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    rng = np.random.default_rng(23)
    cases = [(float(x), float(y)) for x in np.concatenate([rng.uniform(0.01, 40.0, 300), 10.0 ** rng.uniform(-8, 8, 100)])
             for y in (0.33, -0.33, 0.25, 0.75, -0.75, 1.5, -1.5, 2.4, 0.1, 3.9, -7.2)]
    out = subprocess.run([exe], input="".join(f"{x.hex()} {y.hex()}\n" for x, y in cases), capture_output=True, text=True).stdout.split()
    L = PU.O.lib()
    assert len(out) == len(cases)
    for (x, y), g in zip(cases, out):
        assert float.fromhex(g) == L.orc_go_pow(x, y), (x, y, float.fromhex(g), L.orc_go_pow(x, y))
        assert abs(float.fromhex(g) - x ** y) <= 4e-15 * x ** y            # and it IS a power: (|yf ln x| + 2) ulp from the exact one


@pytest.mark.gpu
def test_device_gomath_matches_the_oracle_bit_for_bit(tmp_path):
    exe = _build(tmp_path)
    L = PU.O.lib()
    rng = np.random.default_rng(17)
    nan, inf = float("nan"), float("inf")
    cases = []
    special = [0.0, -0.0, 1.0, -1.0, 2.5, -2.5, inf, -inf, nan, 1e-300, 1e300]
    for fn in ("max", "min"):
        cases += [(fn, a, b) for a in special for b in special]
        cases += [(fn, float(a), float(b)) for a, b in rng.normal(size=(200, 2))]
    cases += [("sin", float(v), 0.0) for v in np.concatenate([rng.uniform(-50, 50, 400), rng.uniform(-1e6, 1e6, 200), [0.0, -0.0, 1e-310, 5e8]])]
    xs = np.concatenate([rng.uniform(0.05, 30.0, 150), [1.0, 2.0, 0.5]])
    for y in (-2.0, 2.0, 3.0, -3.0, 0.5, -0.5, 7.0, 0.0, 1.0, 16.0):
        cases += [("pow", float(x), y) for x in xs]
    cases += [("pow", 0.0, -2.0), ("pow", inf, -2.0), ("pow", -2.0, 3.0), ("pow", -2.0, 2.0)]
    inp = "".join(f"{fn} {float(a).hex() if not math.isnan(a) else 'nan'} {float(b).hex() if not math.isnan(b) else 'nan'}\n" for fn, a, b in cases)
    r = subprocess.run([exe], input=inp, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [float.fromhex(t) if t not in ("nan", "-nan") else nan for t in r.stdout.split()]
    assert len(got) == len(cases)
    ref_fn = {"max": L.orc_go_max, "min": L.orc_go_min, "pow": L.orc_go_pow, "sin": lambda a, b: L.orc_go_sin(a)}
    for (fn, a, b), g in zip(cases, got):
        ref = ref_fn[fn](a, b)
        if math.isnan(ref):
            assert math.isnan(g), (fn, a, b, g)
        else:
            assert g == ref and math.copysign(1.0, g) == math.copysign(1.0, ref), (fn, a, b, g, ref)
