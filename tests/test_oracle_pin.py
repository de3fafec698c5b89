"""The C++ oracle held to an INDEPENDENT restatement of the reference's device + analysis layers
(oracle/pin_numpy.py: written from the Go sources, plain Python floats, numpy.linalg.solve — no code shared with
oracle/engine.hpp or the CUDA device code).  One pin per device kind and per analysis:

  R, C, L, V (DC / SIN / PULSE / PWL), I   rr, rc, rl, rlc, isin, ipulse, ipwl, vpulse, vpwl, idc
  D                                         diode1-5 (OP / DC sweep / transient incl. transit-time charge)
  Q (NPN / PNP)                             bjt1 (OP), bjt3 (OP, transient), pnp_op (finite OP), bjt2 / pnp_tran before the overflow
  M (Level 1 / 2 / 3, NMOS / PMOS, body bias, junction capacitances)   mosfet1 and the tests/extra_decks.py decks
  K, core inductors                         transformer1 / 2 (first 1 500 accepted steps and full run), transformer3
  nested DC sweep                           dio2src, mos_family

Bar: values within 1e-9 relative + 1e-12 at identical stored rows (LAPACK pivots differently from Sparse 1.3: agreement
is limited by kappa*eps, which is why the ill-conditioned K decks are the interesting ones), identical accepted /
rejected step counts, Newton solve counts, failure times.  Decks whose operating point needs Gmin stepping (diode1) and
lanes that overflow to Inf / NaN are outside this pin (LoadGmin acts on the LU's pivot positions; NaN classes depend on
the elimination order).  The parity label stays "unpinned" until vectors of the real Go solver exist (tests/golden/go):
what this removes is the risk of a transcription slip shared by the oracle and the device code."""
import numpy as np
import pytest

import parity_util as PU
from extra_decks import EXTRA
from oracle import pin_numpy as P

T, O = PU.T, PU.O
B = T.BUNDLED


def _close(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(np.isfinite(a), np.isfinite(b)), what
    fin = np.isfinite(a)
    err = np.abs(a - b)[fin]
    tol = PU.RELTOL * np.abs(b)[fin] + PU.ABSTOL
    assert np.all(err <= tol), (what, float((err / tol).max()))


def _tran(text, max_acc=None, tran=None):
    oc = O.OracleCircuit(text)
    t = dict(oc.netlist.tran) if oc.netlist.analysis == 1 else dict(tstart=0.0, tstop=150e-6, tstep=1e-6, tmax=0.0, uic=False)
    if tran:
        t.update(tran)
    tr = P.Transient(P.Circuit(oc.plan), t["tstart"], t["tstop"], t["tstep"], t["tmax"], t["uic"], max_accepted=max_acc)
    rows = tr.run()
    r = oc.run(1, analysis=1, tran=t)
    return tr, rows, r


FULL_TRAN = ["rr", "rc", "isin", "ipulse", "ipwl", "vpulse", "vpwl", "diode2", "diode4", "mosfet1", "transformer3", "rlc", "transformer1"]
EXTRA_TRAN = ["mos2n", "mos3n", "mos1p", "mos2p", "mos3p", "mos2body", "mos1caps", "pnp_small", "bjt3_tran"]


@pytest.mark.parametrize("name", FULL_TRAN + EXTRA_TRAN)
def test_transient_matches_independent_restatement(name):
    """Whole runs: every stored row, accepted / rejected steps, transient and operating-point solve counts, and — for the
    decks whose Newton iteration fails (PMOS Level 2 / 3, pnp_small) — the failure time."""
    text = B[name] if name in B else EXTRA[name][0]
    tr, rows, r = _tran(text)
    w = r["wave"][0, :int(r["n_rows"][0])]
    _close(rows, w, name)
    assert [tr.accepted, tr.rejected, tr.tran_solves, tr.op_solves] == r["counters"][0, :4].tolist()
    if tr.failed_at is None:
        assert int(r["status"][0]) == 0
    else:
        assert int(r["status"][0]) == 2 and tr.failed_at == r["counters"][0, 5:6].view(np.float64)[0]


@pytest.mark.parametrize("name,n_acc", [("rl", 2500), ("transformer2", 2500), ("bjt2", 20), ("pnp_tran", 2)])
def test_transient_prefix_matches_independent_restatement(name, n_acc):
    """The first accepted steps of the remaining decks: rl / transformer2 (full runs take ~6 s each in pure Python; rlc and
    transformer1 are run in full above), and the BJT decks up to the step before the base edge, where the reference
    overflows to Inf / NaN."""
    text = B[name] if name in B else EXTRA[name][0]
    tr, rows, r = _tran(text, max_acc=n_acc)
    assert len(rows) == n_acc
    _close(rows, r["wave"][0, :n_acc], name)


@pytest.mark.parametrize("name", ["diode5", "idc", "bjt1", "pnp_op", "mos2n", "mos3n", "mos1p", "mos2p", "mos3p", "mos2body", "mos1caps",
                                  "pnp_tran", "bjt3_tran", "dio2src"])
def test_operating_point_matches_independent_restatement(name):
    text = B[name] if name in B else EXTRA[name][0]
    oc = O.OracleCircuit(text)
    ck = P.Circuit(oc.plan)
    op = P.OperatingPoint(ck)
    op.Execute()
    r = oc.run(1, analysis=0)
    assert int(r["status"][0]) == 0 and int(r["counters"][0, 4]) == 0
    if name != "bjt1":                       # bjt1's "converged" point is NaN on both sides (NaN passes the test)
        _close(op.result[None, :], r["wave"][0, :1, :oc.n], name)
    assert ck.M.n_solves == int(r["counters"][0, 3])


DC_CASES = [("diode3", None, ("Vin", -1.0, 3.0, 0.1), None)]
for _n in ("mos2n", "mos3n", "mos1p", "mos2p", "mos3p", "mos2body", "dio2src", "mos_family"):
    DC_CASES.append((_n, EXTRA[_n][1].get("dc_text"), EXTRA[_n][1]["dc"], None))
for _n in ("dio2src", "mos_family"):
    DC_CASES.append((_n, None, None, EXTRA[_n][1]["dc2"]))


@pytest.mark.parametrize("name,dc_text,dc,dc2", DC_CASES, ids=[f"{c[0]}-{'nested' if c[3] else 'single'}" for c in DC_CASES])
def test_dc_sweep_matches_independent_restatement(name, dc_text, dc, dc2):
    """Single and nested sweeps (dc.go:88-140, 205-288) incl. the decks whose sweep stops at a non-converging point."""
    text = dc_text or (B[name] if name in B else EXTRA[name][0])
    oc = O.OracleCircuit(text)
    ck = P.Circuit(oc.plan)
    if dc2:
        (s1, a1, b1, c1), (s2, a2, b2, c2) = dc2
        d = P.DCSweep(ck, [s1, s2], [a1, a2], [b1, b2], [c1, c2])
        r = oc.run(1, analysis=3, dc=dict(source=s1, start=a1, stop=b1, inc=c1), dc2=dict(source=s2, start=a2, stop=b2, inc=c2))
    else:
        d = P.DCSweep(ck, [dc[0]], [dc[1]], [dc[2]], [dc[3]])
        r = oc.run(1, analysis=3, dc=dict(source=dc[0], start=dc[1], stop=dc[2], inc=dc[3]))
    rows = d.run()
    _close(rows, r["wave"][0, :int(r["n_rows"][0])], name)
    assert ck.M.n_solves == int(r["counters"][0, 3])
    assert (d.failed_at is None) == (int(r["status"][0]) == 0)
    if d.failed_at is not None:
        assert d.failed_at[0] == r["counters"][0, 5:6].view(np.float64)[0]


def test_pin_with_parameter_overrides():
    """The pin follows a parameter draw like the oracle does (transformer1: L, k and R of one §8(d) draw)."""
    text = B["transformer1"]
    oc = O.OracleCircuit(text)
    ov = PU.draws("transformer1", T.Circuit.from_netlist(text), 3)
    one = {k: float(v[2]) for k, v in ov.items()}
    tr = P.Transient(P.Circuit(oc.plan, one), 0.0, 3e-3, 1e-5, 0.0, False, max_accepted=1200)
    rows = tr.run()
    r = oc.run(1, overrides={k: np.array([v]) for k, v in one.items()})
    _close(rows, r["wave"][0, :1200], "transformer1 draw")
