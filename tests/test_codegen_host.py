"""The generated per-netlist code, compiled for the HOST: the condensed transient elimination of the fast build
(Ckt::prefactor + Ckt::assemble_solve_tf, codegen.cpp: emit_tranfast — invariant pivots first, their part computed once
per instance) must solve the same systems as the reference-order elimination (Ckt::assemble_solve) for arbitrary
parameters, device state, time and step — checked here with g++ on the very text the GPU kernels are compiled from
(struct Ckt is plain C++ over device/models.cuh), for the bundled decks, the extra decks and random decks."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import parity_util as PU
from extra_decks import EXTRA
from random_decks import random_active_deck, random_deck

T = PU.T
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <cstdio>
#include <cstdlib>
#include <cmath>
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
#define __ldcs(p) (*(p))
struct double2 { double x, y; };
#include "MODELS"
struct TsbArgs { long long n_inst; const double* pv[128]; const double* U; double Uc[32]; };
STRUCT
int main(int argc, char** argv) {
    const int trials = 200;
    unsigned long long rng = 88172645463325252ULL;
    auto uni = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (double)(rng >> 11) / 9007199254740992.0; };
    double worst = 0;
    int fails = 0, both_fail = 0;
    for (int t = 0; t < trials; ++t) {
        double U[NPAR + 1], V[128][1];
        TsbArgs a; a.n_inst = 1; a.U = U;
        const double nominal[] = {NOMINAL};
        for (int k = 0; k < NPAR; ++k) { U[k] = nominal[k]; if (k < 32) a.Uc[k] = U[k]; }
        for (int s = 0; s < NVAR; ++s) { const int k = VARIDX[s]; V[s][0] = nominal[k] * std::exp((uni() - 0.5) * 1.4); a.pv[s] = V[s]; }
        Ckt c1, c2;
        c1.load(a, 0); c1.init(); c2.load(a, 0); c2.init();
        c2.prefactor();
        for (int k = 0; k < NSTATE; ++k) { double v = (uni() - 0.5) * (NONLINEAR ? 1.2 : 6.0); c1.S[k] = v; c2.S[k] = v; }
        const double time = uni() * 2e-3, dt = std::exp(std::log(1e-9) + uni() * std::log(1e5));
        c1.eval_sources(time, 1.0); c2.eval_sources(time, 1.0);
        const double rdt = 1.0 / dt;
        bool ok1 = c1.assemble_solve<TSB_MODE_TRAN>(TSB_MODE_TRAN, time, dt, rdt, 0.0);
        bool ok2 = c2.assemble_solve_tf<true>(time, dt, rdt, Ckt::TsbNoMid());
        if (!ok1 || !ok2) { if (!ok1 && !ok2) ++both_fail; else ++fails; continue; }
        double scale = 0;
        for (int i = 1; i <= Ckt::N; ++i) scale = std::fmax(scale, std::fabs(c1.x[i]));
        for (int i = 1; i <= Ckt::N; ++i) {
            if (!std::isfinite(c1.x[i]) || !std::isfinite(c2.x[i])) { if (std::isfinite(c1.x[i]) != std::isfinite(c2.x[i])) ++fails; continue; }
            worst = std::fmax(worst, std::fabs(c1.x[i] - c2.x[i]) / (std::fabs(c1.x[i]) + 1e-6 * scale + 1e-300));
        }
        // device state after the two stamps must agree too (nonlinear stamps have side effects on S)
        for (int k = 0; k < NSTATE; ++k) if (c1.S[k] != c2.S[k] && !(c1.S[k] != c1.S[k])) ++fails;
    }
    printf("worst %.3e fails %d both_fail %d\n", worst, fails, both_fail);
    return 0;
}
'''


def _host_check(text, tmp, fast_tol=1e-7):
    ckt = T.Circuit.from_netlist(text)
    b = ckt.batch(2)
    ov = PU.draws("x", ckt, 2, seed=5)
    for (d, p), v in ov.items():
        b.set_param(d, p, v)
    src = b.kernel_source(T.default_opts(strict_fp=0, min_blocks=2))
    if "HAS_TF = true" not in src:
        return None
    m = re.search(r"struct Ckt \{.*?\n\};\n", src, re.S)
    struct = m.group(0)
    # flat parameter table: nominal values, which indices vary (in slot order)
    devs = ckt.devices()
    nominal, varidx = [], {}
    off = 0
    for d in devs:
        for j, v in enumerate(d["p"]):
            nominal.append(v)
        for (dn, p) in ov:
            if dn == d["name"]:
                varidx[(dn, p)] = off + p
        off += len(d["p"])
    slots = [int(x) for x in re.findall(r"P\[(\d+)\] = __ldcs\(a\.pv\[\d+\]", struct)]
    nstate = int(re.search(r"double S\[(\d+)\]", struct).group(1))
    code = (HARNESS.replace("MODELS", os.path.join(ROOT, "toy-spice_b200", "csrc", "device", "models.cuh")).replace("STRUCT", struct)
            .replace("NPAR", str(len(nominal))).replace("NOMINAL", ", ".join(repr(float(v)) for v in nominal) or "0")
            .replace("NVAR", str(len(slots))).replace("VARIDX", "((const int[]){" + ", ".join(map(str, slots or [0])) + "})")
            .replace("NSTATE", str(nstate)).replace("NONLINEAR", "1" if "HAS_NL = true" in struct else "0"))
    cu = os.path.join(tmp, "h.cpp")
    open(cu, "w").write(code)
    exe = os.path.join(tmp, "h")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-DTSB_FAST_DIV", "-w", "-o", exe, cu], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True).stdout
    worst, fails, both = re.match(r"worst (\S+) fails (\d+) both_fail (\d+)", out).groups()
    return float(worst), int(fails), int(both)


DECKS = {n: T.BUNDLED[n] for n in ("rc", "rl", "rlc", "rr", "isin", "ipulse", "vpwl", "diode2", "diode4", "mosfet1", "transformer1", "transformer2",
                                   "transformer3")}
DECKS.update({n: EXTRA[n][0] for n in ("mos2n", "mos3n", "mos1p", "mos2body", "mos1caps", "dio2src")})
for _s in range(8):
    DECKS[f"random{_s}"] = random_deck(_s)[0]
for _s in (1, 2, 3, 5, 6, 7):
    DECKS[f"active{_s}"] = random_active_deck(_s)[0]


@pytest.mark.parametrize("name", sorted(DECKS))
def test_condensed_transient_elimination_equals_reference_order(built, name):
    with tempfile.TemporaryDirectory() as tmp:
        res = _host_check(DECKS[name], tmp)
    if res is None:
        pytest.skip("no condensed elimination for this deck (BJT / nothing invariant)")
    worst, fails, both = res
    assert fails == 0, res
    # two elimination orders of the same well-posed system: agreement to rounding x condition number
    print(name, res)
    assert worst < (1e-5 if name.startswith(("transformer", "active")) else 1e-7), res
