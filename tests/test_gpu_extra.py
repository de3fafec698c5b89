"""GPU-vs-oracle parity on the device-model paths no bundled deck reaches (tests/extra_decks.py): MOSFET Level 2 / 3
(mosfet.go:378-459) and their finite-difference conductances (:505-533), PMOS, body bias, junction capacitances, PNP
(bjt.go:214-313), bjt1 / bjt3 with the supplied transient card (BASELINE configs[3]), and the nested DC sweep
(dc.go:205-288).  Every deck runs as OP, DC sweep and transient where it has them, in the auto, strict (reference
rounding) and fast kernel builds, on a 24-instance SURVEY §8(d) parameter sweep, through the C ABI."""
import numpy as np
import pytest

import parity_util as PU
from extra_decks import EXTRA

T, O = PU.T, PU.O
pytestmark = pytest.mark.gpu
N = 24

CASES = []
for _name, (_text, _an) in EXTRA.items():
    if _an.get("op"):
        CASES.append((_name, "op"))
    if "tran" in _an:
        CASES.append((_name, "tran"))
    if "dc" in _an:
        CASES.append((_name, "dc"))
    if "dc2" in _an:
        CASES.append((_name, "dc2"))


def _run_both(ctx, name, kind, mode, n=N):
    text, an = EXTRA[name]
    kw = {}
    if kind == "op":
        kw = dict(analysis=T.AN_OP)
    elif kind == "tran":
        kw = dict(analysis=T.AN_TRAN)
    elif kind == "dc":
        text = an.get("dc_text", text)
        kw = dict(analysis=T.AN_DC, dc=an["dc"])
    else:
        kw = dict(analysis=T.AN_DC, dc2=an["dc2"])
    ov = PU.draws(name, T.Circuit.from_netlist(text), n)
    cap = 1 if kind == "op" else 4096
    ckt, batch, ana = PU.run_gpu(ctx, text, n, ov, cap_rows=cap, opts=T.default_opts(strict_fp=mode), **kw)
    okw = dict(kw)
    okw["analysis"] = {T.AN_OP: 0, T.AN_TRAN: 1, T.AN_DC: 3}[kw["analysis"]]
    _, ores = PU.run_oracle(text, n, ov, cap_rows=cap, **okw)
    return batch, ores, ana


@pytest.mark.parametrize("mode", [-1, 1, 0], ids=["auto", "strict", "fast"])
@pytest.mark.parametrize("name,kind", CASES, ids=[f"{n}-{k}" for n, k in CASES])
def test_extra_deck_matches_oracle(ctx, name, kind, mode):
    batch, ores, ana = _run_both(ctx, name, kind, mode)
    # pnp_op's DC sweep overflows on purpose (no junction limiting): Inf-vs-NaN of an overflowed lane hangs on the
    # last bit of exp(); there only finite / non-finite classes are compared (bjt1 / bjt2 / pnp_tran keep the exact classes)
    rep = PU.compare_waves(batch, ores, N, nonfinite_any=(name, kind) == ("pnp_op", "dc"))
    assert PU.report_ok(rep), PU.report_str(rep)
    # discrete decisions (Newton stop iteration): identical up to the documented near-threshold flips
    assert rep["counter_mismatch"] <= 2, PU.report_str(rep)
    # failing lanes fail at the same time / sweep value (tran.go:119, dc.go:128)
    st = batch.status()
    cnt = batch.counters()
    bad = np.nonzero(st != 0)[0]
    if bad.size:
        fa_g = cnt[5, bad].view(np.float64)
        fa_o = ores["counters"][bad, 5].copy().view(np.float64)
        assert np.allclose(fa_g, fa_o, rtol=1e-12, atol=0), (fa_g, fa_o)


def test_nested_sweep_columns_and_layout(ctx):
    """SWEEP1 / SWEEP2 columns of StoreNestedResult (dc.go:272-288): source 1 is the outer loop."""
    text, an = EXTRA["mos_family"]
    ckt = T.Circuit.from_netlist(text, ctx)
    (s1, a1, b1, c1), (s2, a2, b2, c2) = an["dc2"]
    dc = T.NewDCSweep([s1, s2], [a1, a2], [b1, b2], [c1, c2])
    dc.Setup(ckt)
    dc.Execute()
    r = dc.GetResults()
    assert list(r)[:2] == ["SWEEP1", "SWEEP2"] and set(r) == {"SWEEP1", "SWEEP2", "V(1)", "V(2)", "I(VDS)", "I(VGS)"}
    n1 = len(np.arange(a1, b1 + 1e-12, c1)); n2 = len(np.arange(a2, b2 + 1e-12, c2))
    assert len(r["SWEEP1"]) == n1 * n2
    assert np.array_equal(r["SWEEP1"], np.repeat(np.arange(n1) * c1 + a1, n2))
    assert np.array_equal(r["SWEEP2"], np.tile(np.arange(n2) * c2 + a2, n1))
    assert np.array_equal(r["V(2)"], r["SWEEP1"]) and np.array_equal(r["V(1)"], r["SWEEP2"])
    # output characteristics: the drain current grows with VGS at fixed VDS (GetSolution's I(VDS) = -x[branch] = +Id)
    idrain = r["I(VDS)"].reshape(n1, n2)
    assert np.all(np.diff(idrain[:, -1]) >= 0) and idrain[-1, -1] > 1e-4
    # three sources: the reference's error (dc.go:86)
    dc3 = T.NewDCSweep(["VDS", "VGS", "VDS"], [0, 0, 0], [1, 1, 1], [1, 1, 1])
    dc3.Setup(ckt)
    with pytest.raises(T.TsbError, match="unsupported number of sweep sources: 3"):
        dc3.Execute()


def test_nested_sweep_statistics_output(ctx):
    text, an = EXTRA["dio2src"]
    n = 40
    ov = PU.draws("dio2src", T.Circuit.from_netlist(text), n)
    _, bw, _ = PU.run_gpu(ctx, text, n, ov, analysis=T.AN_DC, dc2=an["dc2"], out=T.OUT_WAVE | T.OUT_STATS)
    w = bw.wave_all()
    s = bw.stats_all()
    assert np.array_equal(s[0], w.min(axis=0)) and np.array_equal(s[1], w.max(axis=0)) and np.array_equal(s[3], w[-1])
    assert np.allclose(s[2], w.sum(axis=0), rtol=1e-12, atol=1e-300)
