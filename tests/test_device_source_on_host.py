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
from random_decks import diode_rc_ladder, random_active_deck, random_deck, rc_ladder

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
    _, ores = PU.run_oracle(text, n, ov, want_order_sig=True, want_stats=True, **okw)
    nominal_sig = PU.run_oracle(text, 1, {}, want_wave=False, want_order_sig=True, **okw)[1]["order_sig"][0]
    is_op = kw.get("analysis", ckt0.analysis_card()["analysis"]) == T.AN_OP
    return hb, ores, ores["order_sig"] == nominal_sig, is_op


@pytest.mark.parametrize("name,kind", CASES, ids=[f"{n}-{k}" for n, k in CASES])
def test_device_source_reproduces_the_oracle(built, name, kind):
    n = N_INST
    hb, ores, same_order, is_op = _host_vs_oracle(name, kind, n)
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
            if nr > 0 and not is_op and not np.isnan(wo).any():        # the running statistics of the result store: min, max, sum, last per column
                assert np.array_equal(hb.stats_all()[:, :, i], ores["stats"][i][:, : ores["ncol"]]), (name, i)
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


VARIANTS = {
    # the transient inside the state machine instead of the warp-synchronous loops (tsb_opts.lane_refill: the code path of
    # free-running lanes), the reference's literal second solve per linear step, and the host-compilable part of the fast
    # build (x * (1/dt), hoisted invariants, condensed transient elimination; the device-only reciprocal seeds are not here)
    "lane_refill": dict(opts_kw=dict(lane_refill=1)),
    "literal_second_solve": dict(opts_kw=dict(skip_linear_resolve=0)),
    "fast_build": dict(strict=False),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("name", ["rc", "rlc", "diode1", "diode2", "mosfet1", "bjt2", "transformer1", "mos2n"])
def test_other_code_paths_of_the_device_source(built, name, variant):
    """Same check over the other drivers / builds of the same source: bit identity where the arithmetic is the reference's
    (any mapping of the strict build), the 1e-9 / 1e-12 contract with identical rows, status and step counts for the fast
    build's re-associations."""
    n = N_INST
    text = DECKS[name]
    ov = PU.draws(name, T.Circuit.from_netlist(text), n)
    with tempfile.TemporaryDirectory() as tmp:
        _, hb, _ = H.run(text, n, ov, tmp, **VARIANTS[variant])
    _, ores = PU.run_oracle(text, n, ov, want_order_sig=True)
    same_order = ores["order_sig"] == PU.run_oracle(text, 1, {}, want_wave=False, want_order_sig=True)[1]["order_sig"][0]
    rep = PU.compare_waves(hb, ores, n)
    assert PU.report_ok(rep), PU.report_str(rep)
    assert rep["counter_mismatch"] <= (1 if name == "diode1" else 0), PU.report_str(rep)
    if variant != "fast_build":
        for i in np.nonzero(same_order)[0]:
            nr = int(ores["n_rows"][i])
            assert np.array_equal(hb.waveform(i), ores["wave"][i, :nr, : ores["ncol"]], equal_nan=True), (name, variant, int(i))


@pytest.mark.parametrize("name", ["diode1", "diode2", "mosfet1", "bjt2", "pnp_tran"])
def test_one_lane_refilling_itself_through_the_whole_batch(built, name):
    """Lane refill (tsb_opts.lane_refill, nonlinear circuits): a lane that finishes an instance writes its results, takes the
    next unprocessed instance from the work counter and re-enters the same loop.  Here ONE emulated lane starts on instance 0
    and works through all of them that way — whatever state an instance leaves behind (device state, result store, step
    control, operating-point fallback stage) must not reach the next one."""
    n = 10
    text = DECKS[name]
    ov = PU.draws(name, T.Circuit.from_netlist(text), n)
    with tempfile.TemporaryDirectory() as tmp:
        _, hb, _ = H.run(text, n, ov, tmp, opts_kw=dict(lane_refill=1), refill_chain=True)
    _, ores = PU.run_oracle(text, n, ov, want_order_sig=True)
    same_order = ores["order_sig"] == PU.run_oracle(text, 1, {}, want_wave=False, want_order_sig=True)[1]["order_sig"][0]
    rep = PU.compare_waves(hb, ores, n)
    assert PU.report_ok(rep) and rep["counter_mismatch"] <= (1 if name == "diode1" else 0), PU.report_str(rep)
    for i in np.nonzero(same_order)[0]:
        nr = int(ores["n_rows"][i])
        assert np.array_equal(hb.waveform(i), ores["wave"][i, :nr, : ores["ncol"]], equal_nan=True), (name, int(i))


@pytest.mark.parametrize("name,grid_dt", [("rc", 0.0), ("rlc", 1e-6), ("diode2", 2e-5), ("mosfet1", 0.0)])
def test_fixed_grid_output_of_the_device_source(built, name, grid_dt):
    """TSB_OUT_GRID (interpolated output points under adaptive stepping): the device resamples the series while it runs;
    here its source does so on the host, against the same definition applied in numpy to the oracle's full series."""
    n = 6
    text = T.BUNDLED[name]
    ov = PU.draws(name, T.Circuit.from_netlist(text), n)
    with tempfile.TemporaryDirectory() as tmp:
        _, hb, _ = H.run(text, n, ov, tmp, grid_dt=grid_dt)
    _, ores = PU.run_oracle(text, n, ov, cap_rows=24000)
    tg = hb.grid_times
    assert np.array_equal(hb.status(), ores["status"]) and np.all(hb.rows() == len(tg))
    assert np.array_equal(hb.counters()[7], ores["n_rows"])            # rows of the reference series
    for i in range(n):
        ref = PU.resample_reference(ores["wave"][i], int(ores["n_rows"][i]), ores["ncol"], tg)
        g = hb.wave_all()[: len(tg), :, i]
        assert np.array_equal(g[:, 0], tg)
        scale = np.maximum(np.abs(ref), np.nanmax(np.abs(ref), axis=0, keepdims=True) * 1e-3)
        assert np.all(np.abs(g - ref) <= PU.RELTOL * scale + PU.ABSTOL), (name, i, float(np.max(np.abs(g - ref))))


@pytest.mark.parametrize("name", ["rc", "rl", "rlc", "vpulse", "transformer1"])
def test_shared_time_grid_of_the_device_source(built, name):
    """tsb_opts.share_time_grid: a pilot instance publishes the time-only quantities of its step attempts, the other
    instances verify (time, dt) and reuse them, falling back to their own arithmetic where their step sequence departs from
    the pilot's.  Emulated sequentially (pilot first), the readers must give the reference's bits like any other mapping."""
    n = 8
    text = T.BUNDLED[name]
    ov = PU.draws(name, T.Circuit.from_netlist(text), n)
    with tempfile.TemporaryDirectory() as tmp:
        _, hb, _ = H.run(text, n, ov, tmp, opts_kw=dict(share_time_grid=1))
    assert hb.tgrid_entries is not None and hb.tgrid_entries > 100, hb.tgrid_entries
    _, ores = PU.run_oracle(text, n, ov, want_order_sig=True)
    same_order = ores["order_sig"] == PU.run_oracle(text, 1, {}, want_wave=False, want_order_sig=True)[1]["order_sig"][0]
    rep = PU.compare_waves(hb, ores, n)
    assert PU.report_ok(rep) and rep["counter_mismatch"] == 0, PU.report_str(rep)
    for i in np.nonzero(same_order)[0]:
        nr = int(ores["n_rows"][i])
        assert np.array_equal(hb.waveform(i), ores["wave"][i, :nr, : ores["ncol"]], equal_nan=True), (name, int(i))


RANDOM = {f"random{s}": random_deck(s)[0] for s in range(10)}
RANDOM.update({f"active{s}": random_active_deck(s)[0] for s in range(1, 9)})
RANDOM.update({"rc_ladder12": rc_ladder(12), "diode_ladder8": diode_rc_ladder(8)})


@pytest.mark.parametrize("name", sorted(RANDOM))
def test_device_source_on_random_decks(built, name):
    """Seeded random topologies (passive; with Q / M / K / core inductors; ladders of 10 - 14 unknowns): whatever the code
    generator makes of them is the reference's arithmetic — bit for bit on the instances ordered like the nominal one."""
    n = 8
    text = RANDOM[name]
    ov = PU.draws(name, T.Circuit.from_netlist(text), n)
    with tempfile.TemporaryDirectory() as tmp:
        _, hb, _ = H.run(text, n, ov, tmp)
    _, ores = PU.run_oracle(text, n, ov, want_order_sig=True)
    same_order = ores["order_sig"] == PU.run_oracle(text, 1, {}, want_wave=False, want_order_sig=True)[1]["order_sig"][0]
    rep = PU.compare_waves(hb, ores, n)
    assert PU.report_ok(rep) and rep["counter_mismatch"] == 0, PU.report_str(rep)
    for i in np.nonzero(same_order)[0]:
        nr = int(ores["n_rows"][i])
        assert np.array_equal(hb.waveform(i), ores["wave"][i, :nr, : ores["ncol"]], equal_nan=True), (name, int(i))
        assert np.array_equal(hb.counters()[:4, i], ores["counters"][i, :4]), (name, int(i))


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
