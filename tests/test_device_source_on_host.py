"""The DEVICE source on the host (tests/host_emul.py): models.cuh + skeleton.cuh + the generated per-netlist code — the very
text the sm_100a kernels are compiled from — built with g++ in the strict configuration and run one instance at a time,
whole analyses (operating point with its Gmin / source-stepping fallbacks, transient, DC sweep, nested DC sweep), against
the oracle.

What it establishes, on a machine without a GPU:
* an instance for which the reference's sparse module makes the SAME pivot choices on its own values as on the nominal
  ones is reproduced BIT FOR BIT: stored rows, status, accepted / rejected steps, solve counts, every voltage and current;
* the others (the engine eliminates every instance in the frozen order of the nominal instance — BASELINE north_star:
  "a fixed pivot order taken from the reference's symbolic pass" — where the reference, run on that instance alone, breaks
  a Markowitz tie the other way) differ by the rounding of another elimination order: inside 1e-9 / 1e-12, with identical
  rows and status.  The oracle reports the pivot-choice signature per instance (oracle/sparse13.hpp: sp13_order_sig)."""
import tempfile

import numpy as np
import pytest

import host_emul as H
import parity_util as PU
from extra_decks import EXTRA

T = PU.T
N_INST = 12
DECKS = {n: T.BUNDLED[n] for n in sorted(T.BUNDLED)}
DECKS.update({n: EXTRA[n][0] for n in sorted(EXTRA)})
# lanes that FAIL on both sides iterate a chaotic map to the same failure (PMOS Level 2 / 3: the reference's model does not
# converge): rows and status are compared, counters and values are not
FAILING = {"mos2p", "mos3p"}


# (deck, analysis): every bundled deck with its own analysis card; the extra decks with every analysis tests/extra_decks.py lists
CASES = [(n, "card") for n in sorted(T.BUNDLED)]
for _name, (_text, _an) in sorted(EXTRA.items()):
    CASES += [(_name, k) for k in ("op", "tran", "dc", "dc2") if (_an.get(k) if k != "tran" else "tran" in _an)]


def _host_vs_oracle(name, kind, n):
    text = DECKS[name]
    kw = {}
    if kind != "card":
        an = EXTRA[name][1]
        if kind == "op":
            kw = dict(analysis=T.AN_OP)
        elif kind == "tran":
            kw = dict(analysis=T.AN_TRAN)
        elif kind == "dc":
            text = an.get("dc_text", text)
            kw = dict(analysis=T.AN_DC, dc=an["dc"])
        else:
            kw = dict(analysis=T.AN_DC, dc2=an["dc2"])
    ckt0 = T.Circuit.from_netlist(text)
    if kind == "card" and ckt0.analysis_card()["analysis"] == T.AN_AC:
        pytest.skip(".ac card: the AC path has its own host-compiled check (tests/test_ac.py)")
    ov = PU.draws(name, ckt0, n)
    with tempfile.TemporaryDirectory() as tmp:
        _, hb, _ = H.run(text, n, ov, tmp, **kw)
    okw = dict(kw)
    if "analysis" in okw:
        okw["analysis"] = {T.AN_OP: 0, T.AN_TRAN: 1, T.AN_DC: 3}[okw["analysis"]]
    _, ores = PU.run_oracle(text, n, ov, want_order_sig=True, **okw)
    nominal_sig = PU.run_oracle(text, 1, {}, want_wave=False, want_order_sig=True, **okw)[1]["order_sig"][0]
    return hb, ores, ores["order_sig"] == nominal_sig


@pytest.mark.parametrize("name,kind", CASES, ids=[f"{n}-{k}" for n, k in CASES])
def test_device_source_reproduces_the_oracle(built, name, kind):
    n = N_INST
    hb, ores, same_order = _host_vs_oracle(name, kind, n)
    st, rows, cnt = hb.status(), hb.rows(), hb.counters()
    n_exact = 0
    for i in range(n):
        nr = int(ores["n_rows"][i])
        assert int(st[i]) == int(ores["status"][i]) and int(rows[i]) == nr, (name, i, int(st[i]), int(ores["status"][i]), int(rows[i]), nr)
        if name in FAILING and int(st[i]) != 0:
            continue
        wg, wo = hb.waveform(i), ores["wave"][i, :nr, : ores["ncol"]]
        counters_equal = np.array_equal(cnt[:4, i], ores["counters"][i, :4])
        if (name, kind) == ("pnp_op", "dc"):
            # overflows on purpose (no junction limiting): whether an overflowed entry reads Inf or NaN hangs on
            # which exact zeros the elimination multiplies; finite / non-finite classes and the finite values are compared
            fin = np.isfinite(wo)
            assert np.array_equal(np.isfinite(wg), fin), (name, i)
            assert np.all(np.abs(wg[fin] - wo[fin]) <= 1e-9 * np.abs(wo[fin]) + 1e-12), (name, i)
            n_exact += int(same_order[i])
            continue
        if same_order[i]:
            assert counters_equal, (name, i, cnt[:4, i].tolist(), ores["counters"][i, :4].tolist())
            assert np.array_equal(wg, wo, equal_nan=True), (name, i, float(np.nanmax(np.abs(wg - wo))))
            n_exact += 1
        else:
            # another elimination order: NaN / Inf classes and values inside the contract (a solve COUNT may move by one on an
            # operating point that goes through 100 non-converging iterations before Gmin stepping: diode1 / diode5)
            fin = np.isfinite(wo) & np.isfinite(wg)
            assert np.array_equal(np.isnan(wg), np.isnan(wo)) and np.array_equal(np.isinf(wg), np.isinf(wo)), (name, i)
            assert np.all(np.abs(wg[fin] - wo[fin]) <= 1e-9 * np.abs(wo[fin]) + 1e-12), (name, i, float(np.max(np.abs(wg[fin] - wo[fin]))))
    if name not in FAILING:
        assert n_exact == int(same_order.sum())
    print(name, kind, f"same pivot order as the nominal instance: {int(same_order.sum())}/{n}, all of them bit-identical")


def test_the_reference_orders_some_instances_differently(built):
    """The premise of the split above, pinned: on rlc / diode2 every draw is ordered like the nominal instance; on diode1 and
    transformer2 a good share is not — which is what the 1e-9 / 1e-12 contract (rather than bit equality) is for."""
    share = {}
    for name in ("rlc", "diode2", "diode1", "transformer2"):
        text = T.BUNDLED[name]
        ov = PU.draws(name, T.Circuit.from_netlist(text), 64)
        sig = PU.run_oracle(text, 64, ov, want_wave=False, want_order_sig=True)[1]["order_sig"]
        nominal = PU.run_oracle(text, 1, {}, want_wave=False, want_order_sig=True)[1]["order_sig"][0]
        share[name] = float((sig == nominal).mean())
    assert share["rlc"] == 1.0 and share["diode2"] == 1.0, share
    assert 0.2 < share["diode1"] < 0.9 and 0.5 < share["transformer2"] < 1.0, share
