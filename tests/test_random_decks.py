"""Random decks (tests/random_decks.py): front-end, numbering and symbolic pass of the product against the oracle's
independent restatement on CPU; full analysis parity on the GPU."""
import numpy as np
import pytest

import parity_util as PU
from random_decks import random_active_deck, random_deck

T, O, onl = PU.T, PU.O, PU.onl
SEEDS = list(range(40))


@pytest.mark.parametrize("seed", SEEDS)
def test_front_end_and_symbolic_pass_on_random_decks(built, seed):
    text, _ = random_deck(seed)
    ckt = T.Circuit.from_netlist(text)
    oc = O.OracleCircuit(text)
    assert ckt.GetNodeMap() == oc.plan.node_map and ckt.GetBranchMap() == oc.plan.branch_map
    dv = ckt.devices()
    assert len(dv) == len(oc.plan.devices)
    for a, b in zip(dv, oc.plan.devices):
        assert (a["kind"], a["name"], a["nodes"], a["branch"], a["p"], a["ip"]) == (b.kind, b.name, list(b.nodes), b.branch, list(b.p), list(b.ip)), text
    card, nl = ckt.analysis_card(), oc.netlist
    assert (card["analysis"], card["tstart"], card["tstop"], card["tstep"], card["tmax"]) == (
        nl.analysis, nl.tran["tstart"], nl.tran["tstop"], nl.tran["tstep"], nl.tran["tmax"])
    st, so = ckt.structure(), oc.structure()
    assert st["ext2int"] == so["ext2int"], text
    assert st["pivot_row"] == so["pivot_row"] and st["pivot_col"] == so["pivot_col"], text
    # the generated kernel source exists for every deck (specialisation never refuses a netlist of supported devices)
    assert "tsb_optran" in ckt.batch(2).kernel_source(T.default_opts(min_blocks=2))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [-1, 1], ids=["auto", "strict"])
@pytest.mark.parametrize("seed", SEEDS[:24])
def test_random_deck_matches_oracle(ctx, seed, mode):
    text, info = random_deck(seed)
    n = 6
    ov = PU.draws("random", T.Circuit.from_netlist(text), n, seed=1000 + seed)
    cap = 24000 if info["has_inductor"] else 2048
    ckt, batch, an = PU.run_gpu(ctx, text, n, ov, cap_rows=cap, opts=T.default_opts(strict_fp=mode))
    _, ores = PU.run_oracle(text, n, ov, cap_rows=cap)
    rep = PU.compare_waves(batch, ores, n)
    assert PU.report_ok(rep), (rep, text)
    assert rep["compared_points"] > 0
    assert rep["counter_mismatch"] <= 1, (rep, text)


ACTIVE_SEEDS = list(range(24))


@pytest.mark.parametrize("seed", ACTIVE_SEEDS)
def test_front_end_and_symbolic_pass_on_random_active_decks(built, seed):
    """BJT / MOSFET (all levels, both polarities) / coupled-inductor / core-inductor decks: both front-ends produce the same
    device table (model cards, instance parameters, `core=` inductors, K over 2-3 windings) and both symbolic passes the
    same pivot order."""
    text, _ = random_active_deck(seed)
    ckt = T.Circuit.from_netlist(text)
    oc = O.OracleCircuit(text)
    assert ckt.GetNodeMap() == oc.plan.node_map and ckt.GetBranchMap() == oc.plan.branch_map
    dv = ckt.devices()
    assert len(dv) == len(oc.plan.devices)
    for a, b in zip(dv, oc.plan.devices):
        assert (a["kind"], a["name"], a["nodes"], a["branch"], a["p"], a["ip"]) == (b.kind, b.name, list(b.nodes), b.branch, list(b.p), list(b.ip)), text
    st, so = ckt.structure(), oc.structure()
    assert st["ext2int"] == so["ext2int"], text
    assert st["pivot_row"] == so["pivot_row"] and st["pivot_col"] == so["pivot_col"], text
    assert "tsb_optran" in ckt.batch(2).kernel_source(T.default_opts(min_blocks=2))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [-1, 1], ids=["auto", "strict"])
@pytest.mark.parametrize("seed", ACTIVE_SEEDS[:16])
def test_random_active_deck_matches_oracle(ctx, seed, mode):
    text, info = random_active_deck(seed)
    n = 6
    ov = PU.draws("random", T.Circuit.from_netlist(text), n, seed=2000 + seed)
    cap = 24000 if info["has_inductor"] else 2048
    ckt, batch, an = PU.run_gpu(ctx, text, n, ov, cap_rows=cap, opts=T.default_opts(strict_fp=mode))
    _, ores = PU.run_oracle(text, n, ov, cap_rows=cap)
    rep = PU.compare_waves(batch, ores, n)
    assert PU.report_ok(rep), (rep, text)
    assert rep["counter_mismatch"] <= 1, (rep, text)


@pytest.mark.gpu
@pytest.mark.parametrize("coop", [0, -1])
@pytest.mark.parametrize("sections", [12, 24])
def test_rc_ladder_beyond_the_bundled_sizes(ctx, sections, coop):
    """n = 14 and 26 unknowns: past the point where a circuit fits in one thread's registers (the generated kernel
    spills to local memory) the thread-per-circuit path (coop_parts = 0) must still be CORRECT; the default (-1) runs the
    n = 26 ladder on the cooperative mapping (tests/test_gpu_coop.py has its own parity decks)."""
    from random_decks import rc_ladder
    text = rc_ladder(sections)
    n = 6
    ov = PU.draws("ladder", T.Circuit.from_netlist(text), n, seed=77)
    ckt, batch, an = PU.run_gpu(ctx, text, n, ov, cap_rows=1024, opts=T.default_opts(coop_parts=coop))
    _, ores = PU.run_oracle(text, n, ov, cap_rows=1024)
    rep = PU.compare_waves(batch, ores, n)
    assert PU.report_ok(rep), rep
    assert rep["compared_points"] > 0 and rep["counter_mismatch"] == 0
