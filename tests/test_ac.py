"""AC analysis (ac.go:21-126): the oracle's restatement held to an independent NumPy restatement; the product's
generated per-netlist AC code compiled for the HOST and held to the oracle bit for bit (strict arithmetic); the front-end's
`.ac` card; and — on the GPU — tsb_run_ac through the C ABI against the oracle on seeded parameter draws."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import parity_util as PU
from ac_decks import AC_DECKS, AC_SINGULAR, ac_draws
from oracle import pin_numpy as P

T, O = PU.T, PU.O
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _card(text):
    a = O.OracleCircuit(text).netlist.ac
    return a["sweep"], a["points"], a["fstart"], a["fstop"]


def _close(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = PU.RELTOL * np.abs(b) + PU.ABSTOL
    assert np.all(err <= tol), (what, float((err / tol).max()))


# ------------------------------------------------------------------------------------------------ oracle pins (CPU)
@pytest.mark.parametrize("refread", [False, True])
@pytest.mark.parametrize("name", sorted(AC_DECKS))
def test_oracle_ac_matches_independent_restatement(name, refread):
    oc = O.OracleCircuit(AC_DECKS[name])
    r = oc.run(ac_refread=refread)
    assert r["status"][0] == 0 and r["n_rows"][0] == oc.netlist.ac["points"]
    rows, fail = P.ac_sweep(oc.plan, *_card(AC_DECKS[name]), refread=refread)
    assert fail is None
    _close(r["wave"][0, :len(rows)], rows, name)
    assert r["signals"][0] == "FREQ" and r["signals"][1].endswith("_MAG") and r["signals"][2].endswith("_PHASE")


def test_oracle_ac_low_pass_is_the_textbook_transfer_function():
    oc = O.OracleCircuit(AC_DECKS["ac_lowpass"])
    w = oc.run()["wave"][0]
    f = w[:, 0]
    assert f[0] == 1.0 and abs(f[-1] - 1e6) < 1e-6 and np.allclose(np.diff(np.log10(f)), 0.2)
    H = 1.0 / (1.0 + 2j * np.pi * f * 1e3 * 1e-6)
    k = oc.signals().index("V(2)_MAG")
    assert np.allclose(w[:, k], np.abs(H), rtol=1e-12) and np.allclose(w[:, k + 1], np.degrees(np.angle(H)), rtol=1e-10)


@pytest.mark.parametrize("name", sorted(AC_SINGULAR))
def test_oracle_ac_inductors_make_the_reference_matrix_singular(name):
    """inductor.go:43-57 leaves the branch row empty; Mutual / MagneticInductor have no AC case in Stamp at all."""
    oc = O.OracleCircuit(AC_SINGULAR[name])
    r = oc.run()
    assert r["status"][0] == 5 and r["n_rows"][0] == 0
    # fails at the first frequency: Pow(10, Log10(10)) as Go computes it (log10 = log2 * (Ln2/Ln10)), 10 to within an ulp or two
    assert abs(np.frombuffer(r["counters"][0, 5].tobytes(), dtype=np.float64)[0] - 10.0) < 1e-13
    assert abs(P.ac_sweep(oc.plan, *_card(AC_SINGULAR[name]))[1] - 10.0) < 1e-13


def test_oracle_refuses_nonlinear_ac():
    with pytest.raises(RuntimeError):
        O.OracleCircuit(T.BUNDLED["bjt3"]).run()


# ------------------------------------------------------------------------------------------------ front-end (CPU, host library)
def test_ac_card_both_front_ends(built):
    for name, text in {**AC_DECKS, **AC_SINGULAR, "bjt3": T.BUNDLED["bjt3"]}.items():
        card = T.Circuit.from_netlist(text).analysis_card()
        a = O.OracleCircuit(text).netlist.ac
        assert card["analysis"] == T.AN_AC
        assert (card["ac_sweep"], card["ac_points"], card["ac_fstart"], card["ac_fstop"]) == (a["sweep"], a["points"], a["fstart"], a["fstop"]), name
    ckt = T.Circuit.from_netlist(AC_DECKS["ac_ladder"])
    assert ckt.columns(T.AN_AC) == O.OracleCircuit(AC_DECKS["ac_ladder"]).signals()
    v1 = [d for d in ckt.devices() if d["name"] == "V1"][0]
    assert list(v1["p"]) == [0.0, 2.0, 30.0]                      # DC 0 in OP / tran; magnitude and phase ride along
    an = T.analysis_from_card(ckt)
    assert isinstance(an, T.ACAnalysis) and (an.pointsType, an.numPoints) == ("OCT", 40)


# ------------------------------------------------------------------------------------------------ generated code on the host
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
    double U[NPAR + 1], V[128][1];
    TsbArgs a; a.n_inst = 1; a.U = U;
    const double nominal[] = {NOMINAL};
    for (int k = 0; k < NPAR; ++k) { U[k] = nominal[k]; if (k < 32) a.Uc[k] = U[k]; }
    const int nvar = NVAR, nf = atoi(argv[1]), refread = atoi(argv[2]);
    for (int s = 0; s < nvar; ++s) { V[s][0] = strtod(argv[3 + s], nullptr); a.pv[s] = V[s]; }
    Ckt c;
    c.load(a, 0); c.init();
    for (int k = 0; k < nf; ++k) {
        const double freq = strtod(argv[3 + nvar + k], nullptr), omega = 6.283185307179586 * freq;
        double xr[Ckt::N + 1], xi[Ckt::N + 1], row[Ckt::NCOL_AC];
        if (!c.solve_ac(omega, xr, xi)) { printf("FAIL %a\n", freq); return 0; }
        row[0] = freq;
        c.signals_ac(xr, xi, refread != 0, row + 1);
        for (int j = 0; j < Ckt::NCOL_AC; ++j) printf("%a ", row[j]);
        printf("\n");
    }
    return 0;
}
'''


def _host_ac(text, tmp, ov, freqs, refread):
    ckt = T.Circuit.from_netlist(text)
    b = ckt.batch(2)
    for (d, p), v in ov.items():
        b.set_param(d, p, v)
    src = b.kernel_source(T.default_opts(strict_fp=1, min_blocks=2))
    struct = re.search(r"struct Ckt \{.*?\n\};\n", src, re.S).group(0)
    nominal, owner = [], []
    for d in ckt.devices():
        for j, v in enumerate(d["p"]):
            nominal.append(v)
            owner.append((d["name"], j))
    slots = [int(x) for x in re.findall(r"P\[(\d+)\] = __ldcs\(a\.pv\[\d+\]", struct)]
    code = (HARNESS.replace("MODELS", os.path.join(ROOT, "toy-spice_b200", "csrc", "device", "models.cuh")).replace("STRUCT", struct)
            .replace("NPAR", str(len(nominal))).replace("NOMINAL", ", ".join(repr(float(v)) for v in nominal) or "0")
            .replace("NVAR", str(len(slots))))
    cpp, exe = os.path.join(tmp, "h.cpp"), os.path.join(tmp, "h")
    open(cpp, "w").write(code)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-o", exe, cpp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    args = [str(len(freqs)), str(int(refread))] + [float(ov[owner[k]][0]).hex() for k in slots] + [float(f).hex() for f in freqs]
    out = subprocess.run([exe] + args, capture_output=True, text=True).stdout.strip().splitlines()
    return out


@pytest.mark.parametrize("refread", [False, True])
@pytest.mark.parametrize("name", sorted(AC_DECKS))
def test_generated_ac_code_on_the_host_equals_the_oracle(built, name, refread):
    """struct Ckt's solve_ac / signals_ac (the text the GPU kernels are compiled from) with g++ -ffp-contract=off.  With the
    netlist's own values the product's frozen pivot order IS the oracle's: the same complex elimination, operation for
    operation — magnitudes bit for bit, phases to an ulp of atan2.  With drawn parameters the oracle orders each instance
    for itself (the product keeps the nominal instance's order): agreement to rounding."""
    text = AC_DECKS[name]
    ckt = T.Circuit.from_netlist(text)
    drawn = PU.draws(name, ckt, 2, seed=11)
    nominal = {k: np.full(2, [d for d in ckt.devices() if d["name"] == k[0]][0]["p"][k[1]]) for k in drawn}
    for ov, exact in ((nominal, True), (drawn, False)):
        r = O.OracleCircuit(text).run(2, overrides=ov, ac_refread=refread)
        w = r["wave"][0, :int(r["n_rows"][0])]
        with tempfile.TemporaryDirectory() as tmp:
            out = _host_ac(text, tmp, ov, w[:, 0], refread)
        got = np.array([[float.fromhex(x) for x in line.split()] for line in out])
        assert got.shape == w.shape and np.array_equal(got[:, 0], w[:, 0])
        if exact:
            assert np.array_equal(got[:, 1::2], w[:, 1::2]), name
            assert np.allclose(got[:, 2::2], w[:, 2::2], rtol=4e-16, atol=1e-300), name
        else:
            assert np.allclose(got, w, rtol=1e-13, atol=1e-15), name


@pytest.mark.parametrize("name", sorted(AC_SINGULAR))
def test_generated_ac_code_reports_the_singular_matrix(built, name):
    text = AC_SINGULAR[name]
    ckt = T.Circuit.from_netlist(text)
    ov = PU.draws(name, ckt, 2, seed=11)
    with tempfile.TemporaryDirectory() as tmp:
        out = _host_ac(text, tmp, ov, [10.0, 100.0], False)
    assert len(out) == 1 and out[0].startswith("FAIL ") and float.fromhex(out[0].split()[1]) == 10.0


# ------------------------------------------------------------------------------------------------ GPU, through the C ABI
def _gpu_ac(ctx, text, n, ov, out, refread=False, card=None):
    ckt = T.Circuit.from_netlist(text, ctx)
    b = ckt.batch(n)
    for (d, p), v in ov.items():
        b.set_param(d, p, v)
    sweep, pts, f0, f1 = card or _card(text)
    an = T.NewAC(f0, f1, pts, sweep)
    an.out = out | (T.OUT_AC_REFREAD if refread else 0)
    an.Setup(b)
    an.Execute()
    return ckt, b, an


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [1, 0])
@pytest.mark.parametrize("name", sorted(AC_DECKS))
def test_gpu_ac_matches_oracle(ctx, name, strict):
    n = 48
    text = AC_DECKS[name]
    ov = ac_draws(T.Circuit.from_netlist(text).devices(), n)         # R / C draws + magnitude and phase of the AC source
    for refread in (False, True):
        ckt = T.Circuit.from_netlist(text, ctx)
        b = ckt.batch(n)
        for (d, p), v in ov.items():
            b.set_param(d, p, v)
        sweep, pts, f0, f1 = _card(text)
        b.run_ac(sweep, pts, f0, f1, out=T.OUT_WAVE | T.OUT_STATS | (T.OUT_AC_REFREAD if refread else 0), opts=T.default_opts(strict_fp=strict))
        b.sync()
        _, ores = PU.run_oracle(text, n, ov, want_stats=True, ac_refread=refread)
        rep = PU.compare_waves(b, ores, n)
        assert PU.report_ok(rep) and rep["compared_points"] == n * pts * ores["ncol"], PU.report_str(rep)
        wg = np.transpose(b.wave_all(), (2, 0, 1))
        assert np.allclose(wg, ores["wave"][:, :pts], rtol=1e-10, atol=1e-13)        # in fact rounding-level agreement
        s = b.stats_all()
        so = np.transpose(ores["stats"], (1, 2, 0))
        assert np.allclose(s[[0, 1, 3]], so[[0, 1, 3]], rtol=1e-9, atol=1e-12) and np.allclose(s[2], so[2], rtol=1e-9, atol=1e-9)
        assert ckt.columns(T.AN_AC) == ores["signals"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(AC_SINGULAR))
def test_gpu_ac_singular_like_the_reference(ctx, name):
    n = 40
    text = AC_SINGULAR[name]
    ov = ac_draws(T.Circuit.from_netlist(text).devices(), n)
    _, b, _ = _gpu_ac(ctx, text, n, ov, T.OUT_WAVE)
    assert np.all(b.status() == T.ST_AC_FAILED) and np.all(b.rows() == 0)
    assert np.all(np.abs(b.counters()[5].view(np.float64) - 10.0) < 1e-13)


@pytest.mark.gpu
def test_gpu_ac_refuses_nonlinear_circuits_and_the_api_shape(ctx):
    ckt = T.Circuit.from_netlist(T.BUNDLED["bjt3"], ctx)
    an = T.analysis_from_card(ckt)
    an.Setup(ckt.batch(4))
    with pytest.raises(T.TsbError, match="nonlinear"):
        an.Execute()
    # reference-shaped use: GetResults() keys as StoreACResult names them
    ckt = T.Circuit.from_netlist(AC_DECKS["ac_lowpass"], ctx)
    an = T.analysis_from_card(ckt)
    an.Setup(ckt.batch(1))
    assert an.Execute() is None
    res = an.GetResults()
    assert set(res) == {"FREQ", "V(1)_MAG", "V(1)_PHASE", "V(2)_MAG", "V(2)_PHASE", "I(V1)_MAG", "I(V1)_PHASE"}
    f = np.asarray(res["FREQ"])
    assert len(f) == 31 and np.allclose(res["V(2)_MAG"], np.abs(1 / (1 + 2j * np.pi * f * 1e-3)), rtol=1e-12)
