"""Pins for the CPU oracle (oracle/).  The reference has no tests or golden vectors and cannot be run
here (no Go), so the oracle is pinned by: the one analytic answer the reference implies (rr divider),
the hand-derived structure tables of SURVEY.md Appendix A, the step-count facts of SURVEY.md §6, an
independent NumPy restatement of the reference's rc / rl recurrences, and its own committed
regression vectors (tests/golden/golden.npz)."""
import math
import os

import numpy as np
import pytest

import parity_util as PU

T, O, onl = PU.T, PU.O, PU.onl
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz")


def test_go_sin_matches_libm_within_2ulp():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-20, 20, 4000), rng.uniform(-1e-3, 1e-3, 200), [0.0, math.pi, 2 * math.pi * 1000 * 3e-3]])
    for x in xs:
        a, b = O.go_sin(float(x)), math.sin(float(x))
        assert abs(a - b) <= 2 * np.spacing(max(abs(b), 1e-300)) + 1e-17, (x, a, b)


@pytest.mark.parametrize("v,s", [(2e-7, "200.000 ns"), (1.4e-6, "1.400 us"), (1.999999e-05, "20.000 us"), (1.9994e-05, "19.994 us"), (0.003, "3.000 ms"),
                                 (1.5, "1.500 s"), (2.5e-12, "2.500 ps"), (1e-13, "1.000e-13 s")])
def test_format_value_factor(v, s):          # util/formatter.go:8-24
    assert O.format_value_factor(v) == s


def test_parse_value_quirks():
    assert onl.parse_value("10u") == 9.999999999999999e-06
    assert onl.parse_value("1k") == 1000.0
    assert onl.parse_value("1meg") == 1e6
    assert onl.parse_value("3ms") == 3 * 1e-3
    assert onl.parse_value("5M") == 5.0          # "M" matches the regexp but has no multiplier
    with pytest.raises(onl.NetlistError):
        onl.parse_value("1.")


def test_diode5_model_key_quirk():           # `D (Is=...` stores key "(is" -> Is stays 1e-14
    oc = O.OracleCircuit(T.BUNDLED["diode5"])
    d = next(r for r in oc.plan.devices if r.kind == onl.K_D)
    assert d.p == [1e-14, 1.906, 0.0]


def test_rr_known_answer():
    oc = O.OracleCircuit(T.BUNDLED["rr"])
    op = oc.run(1, analysis=onl.AN_OP)
    assert np.allclose(op["wave"][0, 0, :3], [5.0, 2.5, -2.5e-3], rtol=0, atol=1e-15)
    tr = oc.run(1)
    assert tr["n_rows"][0] == 38 and tr["counters"][0, 0] == 38 and tr["counters"][0, 2] == 76
    w = tr["wave"][0, :38]
    assert tr["signals"] == ["TIME", "V(1)", "V(2)", "I(Vin)", "I(R1)", "I(R2)"]
    assert np.allclose(w[:, 1], 5.0, atol=1e-14) and np.allclose(w[:, 2], 2.5, atol=1e-14)
    assert np.allclose(w[:, 3:], 2.5e-3, atol=1e-17)
    assert w[-1, 0] == 0.003 and w[0, 0] == min(1e-4, 0.003 / 300) / 50.0     # first point at minStep (tran.go:30-37,93)


SIZING = {  # accepted, rejected, transient solves (SURVEY §6 / BASELINE.md §4)
    "rr": (38, 0, 76), "rc": (305, 0, 610), "rl": (20795, 2861, 47312), "rlc": (20795, 2861, 47312),
    "transformer1": (20795, 2861, 47312), "transformer2": (20795, 2861, 47312), "transformer3": (305, 0, 610),
    "diode4": (107, 0, None), "mosfet1": (107, 0, 226), "diode2": (38, 0, None),
}


@pytest.mark.parametrize("name", sorted(SIZING))
def test_step_counts(name):
    res = O.OracleCircuit(T.BUNDLED[name]).run(1, want_wave=False)
    acc, rej, sol = SIZING[name]
    assert res["counters"][0, 0] == acc and res["counters"][0, 1] == rej
    if sol is not None:
        assert res["counters"][0, 2] == sol


# ext -> int Translate numbering, listed as ext2int[1..n]
EXT2INT = {
    "rr": [2, 3, 1], "rc": [2, 3, 1], "diode1": [2, 3, 1], "rl": [2, 3, 1, 4], "rlc": [2, 3, 5, 1, 4],
    "bjt1": [2, 3, 4, 1], "bjt2": [2, 4, 5, 6, 1, 3], "mosfet1": [2, 4, 5, 1, 3],
    "transformer1": [2, 3, 5, 7, 1, 4, 6], "transformer2": [2, 3, 5, 7, 8, 10, 1, 4, 6, 9],
    "transformer3": [2, 3, 5, 6, 1, 4, 7], "vpulse": [2, 1], "idc": [1],
}


@pytest.mark.parametrize("name", sorted(EXT2INT))
def test_translate_numbering(name):
    assert O.OracleCircuit(T.BUNDLED[name]).structure()["ext2int"] == EXT2INT[name]


def test_node_and_branch_numbering():         # circuit.go:48-71, SURVEY §8(a) table
    p = O.OracleCircuit(T.BUNDLED["bjt2"]).plan
    assert p.node_map == {"1": 1, "4": 2, "2": 3, "3": 4} and p.branch_map == {"VCC": 5, "VB": 6}
    p = O.OracleCircuit(T.BUNDLED["transformer2"]).plan
    assert p.branch_map == {"Vin": 7, "Lp": 8, "Ls1": 9, "Ls2": 10}
    p = O.OracleCircuit(T.BUNDLED["diode5"]).plan
    assert p.node_map == {"anode": 1, "n1": 2} and p.branch_map == {"V1": 3}


def _rc_numpy(R, C, amp=5.0, freq=1e3, tstep=1e-5, tstop=3e-3):
    """Independent restatement of what the reference does to rc.cir (no matrix code): V source at the
    START of the step (Q2), capacitor companion with the charge two accepted steps back (Q8), LTE from
    the previous two accepted voltages, step doubling (Q6), StoreTimeResult de-dup (Q22)."""
    tstep = min(tstep, tstop / 300)
    minstep, maxstep = tstep / 50.0, tstep
    V0 = V1 = q0 = q1 = 0.0
    t, dt = 0.0, minstep
    rows = []
    last = None
    while t < tstop:
        nt = t + dt
        if nt > tstop:
            nt = tstop
            dt = nt - t
        vs = 0.0 + amp * O.go_sin(2.0 * math.pi * freq * t + 0.0 * math.pi / 180.0)
        g = 1.0 / R
        geq, ceq = C / dt, q1 / dt
        v2 = (g * vs + ceq) / (g + geq)          # node-2 KCL with v1 = vs
        lte = abs(C * V0 - C * V1) / (2.0 * dt)
        if lte > 7.0 and dt > minstep:
            dt /= 2
            continue
        q1, q0 = q0, C * v2
        V1, V0 = V0, v2
        t = nt
        if last is None or (t != last and O.format_value_factor(t) != O.format_value_factor(last)):
            rows.append((t, vs, v2, (vs - v2) / R))
            last = t
        if t < tstop and dt < maxstep:
            dt = min(dt * 2, maxstep) if lte < 7.0 / 100 else min(dt * 1.1, maxstep)
    return np.array(rows)


@pytest.mark.parametrize("R,C", [(100.0, 1e-6), (57.0, 2.3e-6), (180.0, 0.6e-6)])
def test_rc_against_numpy_restatement(R, C):
    oc = O.OracleCircuit(T.BUNDLED["rc"])
    res = oc.run(1, overrides={("r1", 0): [R], ("c1", 0): [C]})
    ref = _rc_numpy(R, C)
    nr = int(res["n_rows"][0])
    assert nr == len(ref) == 305
    w = res["wave"][0, :nr]
    assert np.array_equal(w[:, 0], ref[:, 0])                       # identical time grid
    assert np.allclose(w[:, 1], ref[:, 1], rtol=1e-13, atol=1e-15)  # V(1) = source
    assert np.allclose(w[:, 2], ref[:, 2], rtol=1e-11, atol=1e-14)  # V(2)
    assert np.allclose(w[:, 4], ref[:, 3], rtol=1e-9, atol=1e-14)   # I(r1)
    assert np.allclose(w[:, 3], w[:, 4], rtol=1e-9, atol=1e-13)     # I(vin) = -x[branch] = I(r1)


def test_inductor_pins_step_near_minstep():
    res = O.OracleCircuit(T.BUNDLED["rl"]).run(1)
    nr = int(res["n_rows"][0])
    t = res["wave"][0, :nr, 0]
    minstep = min(1e-5, 2e-3 / 300) / 50
    d = np.diff(t[:200])
    assert d.max() <= 2.3 * minstep


def test_bjt_goes_nan_like_the_reference_would():     # SURVEY Q13, Q4 (NaN passes convergence)
    res = O.OracleCircuit(T.BUNDLED["bjt2"]).run(1)
    nr = int(res["n_rows"][0])
    assert res["status"][0] == 0 and nr == 156
    assert np.isnan(res["wave"][0, nr - 1, 1:]).all()


def test_golden_vectors_reproduce():
    g = np.load(GOLD)
    decks = sorted({k.split("/")[0] for k in g.files})
    assert len(decks) == 21
    for name in decks:
        keys = [tuple(s.split("|")) for s in g[f"{name}/ov_keys"]]
        ov = {(d, int(p)): g[f"{name}/ov_vals"][i] for i, (d, p) in enumerate(keys)}
        res = O.OracleCircuit(T.BUNDLED[name]).run(4, overrides=ov, cap_rows=12288)
        assert np.array_equal(res["n_rows"], g[f"{name}/n_rows"]), name
        assert np.array_equal(res["status"], g[f"{name}/status"]), name
        assert np.array_equal(res["counters"][:, :5], g[f"{name}/counters"]), name
        for i in range(4):
            idx = g[f"{name}/rows_idx/{i}"]
            assert np.array_equal(res["wave"][i, idx, :res["ncol"]], g[f"{name}/wave/{i}"], equal_nan=True), (name, i)


def test_threaded_runner_is_deterministic():
    oc = O.OracleCircuit(T.BUNDLED["rc"])
    ckt = T.Circuit.from_netlist(T.BUNDLED["rc"])
    ov = PU.draws("rc", ckt, 64)
    a = oc.run(64, overrides=ov, threads=1, cap_rows=320)
    b = oc.run(64, overrides=ov, threads=4, cap_rows=320)
    assert np.array_equal(a["wave"], b["wave"], equal_nan=True) and np.array_equal(a["counters"], b["counters"])


def test_diode_operating_point_satisfies_kirchhoff():
    """Oracle-independent pin of the nonlinear path: at the converged operating point of diode1.cir the resistor current
    equals the diode current of the model equation (diode.go:119-135: Is*(exp(vd/(N*Vt)) - 1), Vt = k*300.15/q), for a
    spread of Is / N draws.  (The reference's Newton test is relative 1e-6 on the solution, so the residual is ~1e-6.)"""
    import parity_util as PU
    T, O = PU.T, PU.O
    text = T.BUNDLED["diode1"]
    n = 64
    ckt = T.Circuit.from_netlist(text)
    ov = PU.draws("diode1", ckt, n)
    res = O.OracleCircuit(text).run(n, overrides=ov)
    names = res["signals"]
    v1 = res["wave"][:, 0, names.index("V(1)")]; v2 = res["wave"][:, 0, names.index("V(2)")]
    Is, N = ov[("d1", 0)], ov[("d1", 1)]
    vt = 1.3806226e-23 * 300.15 / 1.6021918e-19
    i_r = (v1 - v2) / ov[("r1", 0)]
    i_d = Is * (np.exp(v2 / (N * vt)) - 1.0)
    assert np.all(res["status"] == 0)
    assert np.all(np.abs(i_r - i_d) <= 2e-5 * np.abs(i_r)), float(np.max(np.abs(i_r - i_d) / np.abs(i_r)))
    assert np.all((v2 > 0.3) & (v2 < 1.6)) and np.allclose(v1, 5.0)


def test_go_runtime_semantics_restated_in_the_oracle():
    """math.Max / math.Min / math.Pow as the Go specification documents them (src/math/dim.go, pow.go special cases) —
    where C's fmax / fmin / pow differ, the oracle follows Go."""
    import math
    import parity_util as PU
    L = PU.O.lib()
    nan, inf = float("nan"), float("inf")
    # Max / Min: NaN propagates unless an infinity of the winning sign is present; signed zeros are ordered
    assert math.isnan(L.orc_go_max(nan, 1.0)) and math.isnan(L.orc_go_max(1.0, nan)) and math.isnan(L.orc_go_min(nan, 1.0))
    assert L.orc_go_max(inf, nan) == inf and L.orc_go_max(nan, inf) == inf and L.orc_go_min(nan, -inf) == -inf
    assert math.copysign(1.0, L.orc_go_max(0.0, -0.0)) == 1.0 and math.copysign(1.0, L.orc_go_max(-0.0, -0.0)) == -1.0
    assert math.copysign(1.0, L.orc_go_min(0.0, -0.0)) == -1.0 and math.copysign(1.0, L.orc_go_min(0.0, 0.0)) == 1.0
    assert L.orc_go_max(2.0, 3.0) == 3.0 and L.orc_go_min(2.0, 3.0) == 2.0 and L.orc_go_max(-inf, 1.0) == 1.0
    # Pow: documented special cases
    assert L.orc_go_pow(3.7, 0.0) == 1.0 and L.orc_go_pow(1.0, nan) == 1.0 and L.orc_go_pow(nan, 0.0) == 1.0 and L.orc_go_pow(5.5, 1.0) == 5.5
    assert math.isnan(L.orc_go_pow(nan, 2.0)) and math.isnan(L.orc_go_pow(-8.0, 1.0 / 3.0))
    assert L.orc_go_pow(0.0, -2.0) == inf and L.orc_go_pow(-0.0, -3.0) == -inf and L.orc_go_pow(-0.0, 3.0) == 0.0 and L.orc_go_pow(0.0, 2.5) == 0.0
    assert L.orc_go_pow(2.0, inf) == inf and L.orc_go_pow(0.5, inf) == 0.0 and L.orc_go_pow(-1.0, inf) == 1.0 and L.orc_go_pow(2.0, -inf) == 0.0
    assert L.orc_go_pow(inf, -2.0) == 0.0 and L.orc_go_pow(-inf, 3.0) == -inf and L.orc_go_pow(-inf, 2.0) == inf
    assert L.orc_go_pow(2.25, 0.5) == 1.5 and L.orc_go_pow(4.0, -0.5) == 0.5
    # integer exponents: the repeated-squaring sequence — exactly 1 / (x*x) for y = -2 (bjt.go:275), x*x*x for 3
    rng = np.random.default_rng(9)
    for x in rng.uniform(0.1, 50.0, 200):
        assert L.orc_go_pow(float(x), -2.0) == 1.0 / (float(x) * float(x))
        assert L.orc_go_pow(float(x), 3.0) == float(x) * (float(x) * float(x)) or L.orc_go_pow(float(x), 3.0) == (float(x) * float(x)) * float(x)
        assert abs(L.orc_go_pow(float(x), 1.7) - float(x) ** 1.7) <= 4e-16 * float(x) ** 1.7
