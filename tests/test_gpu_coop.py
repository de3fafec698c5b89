"""Cooperative mapping of the transient analysis (tsb_opts.coop_parts; csrc/device/coop.cuh): one instance advanced by 2 or 4
threads in different warps, each eliminating its own sub-circuit.  GPU vs the CPU oracle at the parity contract, and vs the
thread-per-circuit mapping on batches that span several blocks and a partial warp."""
import numpy as np
import pytest

import parity_util as PU
from random_decks import diode_rc_ladder, mos_follower_chain, random_linear_network, rc_ladder, rc_mesh, rlc_ladder

T = PU.T

DECKS = {
    "ladder12": (rc_ladder(12), 1024),
    "ladder24": (rc_ladder(24), 1024),
    "mesh4x5": (rc_mesh(4, 5), 1024),
    "rlcladder4": (rlc_ladder(4), 24000),
    "rlc": (T.BUNDLED["rlc"], 24000),
    "rl": (T.BUNDLED["rl"], 24000),
}
# circuits with nonlinear devices: the Newton-loop form (two barriers per iteration)
DECKS["diodeladder12"] = (diode_rc_ladder(12), 1024)
DECKS["diodeladder24"] = (diode_rc_ladder(24), 1024)
DECKS["moschain12"] = (mos_follower_chain(12), 2048)
for _s in (0, 2, 5, 6):          # random linear networks: irregular graphs, loops, inductor branches, three source waveforms
    _t, _i = random_linear_network(_s, 20)
    DECKS[f"net{_s}"] = (_t, 1024 if _i["inductors"] == 0 else 24000)


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [2, 4, 8])
@pytest.mark.parametrize("name", sorted(DECKS))
def test_cooperative_transient_matches_oracle(ctx, name, parts):
    text, cap = DECKS[name]
    if T.Circuit.from_netlist(text).coop_info(parts) is None:
        pytest.skip(f"no partition into {parts} sub-circuits")
    n = 40 if cap < 2000 else 6          # 40: one full warp and a partial one per part
    ov = PU.draws(name, T.Circuit.from_netlist(text), n, seed=77)
    ckt, batch, an = PU.run_gpu(ctx, text, n, ov, cap_rows=cap, opts=T.default_opts(coop_parts=parts))
    _, ores = PU.run_oracle(text, n, ov, cap_rows=cap)
    rep = PU.compare_waves(batch, ores, n)
    assert PU.report_ok(rep), PU.report_str(rep)
    assert rep["compared_points"] > 0 and rep["counter_mismatch"] == 0, PU.report_str(rep)


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [2, 4, 8])
@pytest.mark.parametrize("name", ["ladder24", "mesh4x5", "rlc", "diodeladder24"])
def test_cooperative_statistics_equal_thread_mapping(ctx, name, parts):
    """Several blocks, a ragged tail, statistics output: same rows / counters / status as the thread-per-circuit mapping, values
    within the contract (a different elimination order is a different rounding)."""
    text, _ = DECKS[name]
    if T.Circuit.from_netlist(text).coop_info(parts) is None:
        pytest.skip(f"no partition into {parts} sub-circuits")
    n = 4096 + 17 if name != "rlc" else 300
    ov = PU.draws(name, T.Circuit.from_netlist(text), n, seed=3)
    _, b0, _ = PU.run_gpu(ctx, text, n, ov, out=T.OUT_STATS, opts=T.default_opts(coop_parts=0, min_blocks=1))     # (the one pre-built variant)
    s0, r0, c0, st0 = b0.stats_all().copy(), b0.rows().copy(), b0.counters().copy(), b0.status().copy()
    _, b1, _ = PU.run_gpu(ctx, text, n, ov, out=T.OUT_STATS, opts=T.default_opts(coop_parts=parts))
    s1, r1, c1, st1 = b1.stats_all(), b1.rows(), b1.counters(), b1.status()
    assert np.array_equal(st0, st1) and np.array_equal(r0, r1)
    assert np.array_equal(c0[:6], c1[:6]) and np.array_equal(c0[7], c1[7])
    # min / max / last at the contract; the sums over ~300 .. 2e4 rows accumulate the per-row differences
    tol = 1e-9 * np.abs(s0) + 1e-12
    assert np.all(np.abs(s1[[0, 1, 3]] - s0[[0, 1, 3]]) <= tol[[0, 1, 3]] * 10)
    assert np.all(np.abs(s1[2] - s0[2]) <= 1e-7 * np.abs(s0[2]) + 1e-9)


@pytest.mark.gpu
def test_cooperative_mapping_refuses_what_it_does_not_cover(ctx):
    text = rc_ladder(12)
    ov = PU.draws("ladder", T.Circuit.from_netlist(text), 4, seed=1)
    with pytest.raises(T.TsbError):
        PU.run_gpu(ctx, text, 4, ov, out=T.OUT_STATS, opts=T.default_opts(coop_parts=4, strict_fp=1))
    with pytest.raises(T.TsbError):
        PU.run_gpu(ctx, text, 4, ov, out=T.OUT_STATS, opts=T.default_opts(coop_parts=3))
    with pytest.raises(T.TsbError):
        PU.run_gpu(ctx, text, 4, ov, out=T.OUT_GRID, opts=T.default_opts(coop_parts=2))
    dtext = T.BUNDLED["bjt2"]            # BJT circuits are generated dense (NaN bookkeeping) and do not partition
    dov = PU.draws("bjt2", T.Circuit.from_netlist(dtext), 4)
    with pytest.raises(T.TsbError):
        PU.run_gpu(ctx, dtext, 4, dov, out=T.OUT_STATS, opts=T.default_opts(coop_parts=2))
    with pytest.raises(T.TsbError):       # rc.cir: three unknowns, nothing to cut
        PU.run_gpu(ctx, T.BUNDLED["rc"], 4, PU.draws("rc", T.Circuit.from_netlist(T.BUNDLED["rc"]), 4), out=T.OUT_STATS, opts=T.default_opts(coop_parts=2))


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [2, 4])
def test_cooperative_uic_and_start_time(ctx, parts):
    """`uic` (no operating point: the hand-over carries zero state) and a start time > 0 (rows before it suppressed, the
    result-store key of those steps never computed) on the cooperative mapping, against the oracle."""
    text = rc_ladder(12)
    n = 12
    ov = PU.draws("ladder", T.Circuit.from_netlist(text), n, seed=9)
    card = T.Circuit.from_netlist(text).analysis_card()
    for tran in ({"uic": True}, {"tstart": 0.4 * card["tstop"]}, {"uic": True, "tstart": 0.25 * card["tstop"]}):
        ckt, batch, an = PU.run_gpu(ctx, text, n, ov, cap_rows=1024, tran=tran, opts=T.default_opts(coop_parts=parts))
        _, ores = PU.run_oracle(text, n, ov, cap_rows=1024, tran=tran)
        rep = PU.compare_waves(batch, ores, n)
        assert PU.report_ok(rep), (tran, PU.report_str(rep))
        assert rep["compared_points"] > 0 and rep["counter_mismatch"] == 0, (tran, PU.report_str(rep))


@pytest.mark.gpu
def test_cooperative_processing_order_changes_no_bit(ctx):
    """tsb_batch_set_order on the cooperative mapping: which instance a group of lanes works on is free; parameters and results
    stay in the caller's order, bit for bit."""
    text = rc_ladder(16)
    n = 1000
    ckt = T.Circuit.from_netlist(text, ctx)
    ov = PU.draws("ladder", ckt, n, seed=4)
    card = ckt.analysis_card()
    res = []
    for perm in (None, np.random.default_rng(1).permutation(n)):
        b = ckt.batch(n)
        for (d, p), v in ov.items():
            b.set_param(d, p, v)
        if perm is not None:
            b.set_order(perm)
        b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS, opts=T.default_opts(coop_parts=2))
        res.append((b.stats_all().copy(), b.rows().copy(), b.counters().copy(), b.status().copy()))
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.gpu
def test_wide_circuit_on_the_thread_mapping(ctx):
    """A 32-section ladder has 67 result columns: the statistics of a 128-thread block (274 KB) exceed an SM's shared memory.
    Thread-per-circuit (coop_parts = 0) must still run — the library shrinks the block to 96 threads — and agree with the
    cooperative mapping, which is what the default picks."""
    text = rc_ladder(32)
    n = 200
    ov = PU.draws("ladder", T.Circuit.from_netlist(text), n, seed=8)
    _, b0, _ = PU.run_gpu(ctx, text, n, ov, out=T.OUT_STATS, opts=T.default_opts(coop_parts=0, min_blocks=1))
    s0, r0, c0 = b0.stats_all().copy(), b0.rows().copy(), b0.counters().copy()
    _, b1, _ = PU.run_gpu(ctx, text, n, ov, out=T.OUT_STATS)
    s1, r1, c1 = b1.stats_all(), b1.rows(), b1.counters()
    assert np.array_equal(r0, r1) and np.array_equal(c0[:6], c1[:6]) and (b0.status() == 0).all() and (b1.status() == 0).all()
    tol = 1e-9 * np.abs(s0) + 1e-12
    assert np.all(np.abs(s1[[0, 1, 3]] - s0[[0, 1, 3]]) <= tol[[0, 1, 3]] * 10)
