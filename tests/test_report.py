"""Result formats (toy-spice_b200/report.py): FormatValueFactor against the oracle's independent restatement, the
reference CLI's table layout on the one deck with an analytic answer, raw-file round trip."""
import ctypes as C
import importlib

import numpy as np

import parity_util as PU

T, O = PU.T, PU.O
R = importlib.import_module("toy-spice_b200.report")


def _oracle_fmt(v):
    buf = C.create_string_buffer(64)
    O.lib().orc_format_value_factor(v, buf, 64)
    return buf.value.decode()


def test_format_value_factor_matches_independent_restatement():
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.normal(size=200) * 10.0 ** rng.integers(-14, 4, 200), [0.0, 1.0, -1.0, 1e-3, 1e-6, 1e-9, 1e-12, 9.9995e-4,
                           0.0025, 0.0035, 2.5e-4, 1234.5675, -0.9999995, 5e-13]])
    for v in vals:
        assert R.format_value_factor(float(v), "s") == _oracle_fmt(float(v)), v


def test_cli_table_layout_rr_operating_point():
    # rr.cir OP: the analytic answer V(1) = 5, V(2) = 2.5, I(Vin) = -2.5 mA (op.go:235-248 does not negate)
    txt = R.format_results({"V(1)": [5.0], "V(2)": [2.5], "I(Vin)": [-2.5e-3]})
    assert txt == ("\nAnalysis Results:\n================\n\nNode Voltages:\nV(1) = 5.000 V\nV(2) = 2.500 V\n"
                   "\nBranch Currents:\nI(Vin) = -2.500 mA\n")


def test_cli_table_layout_transient_and_dc():
    res = {"TIME": [1e-5, 2e-5], "V(2)": [2.5, 2.5], "V(1)": [5.0, 5.0], "I(r1)": [2.5e-3, 2.5e-3], "I(Vin)": [2.5e-3, 2.5e-3]}
    lines = R.format_results(res).split("\n")
    assert lines[4] == "Transient Analysis Results (2 time points):"
    assert lines[5] == "Time        Node Voltages        Branch Currents"
    assert lines[7] == "10.000 us  V(1)=5.000 V  V(2)=2.500 V  I(Vin)=2.500 mA  I(r1)=2.500 mA  "
    dc = R.format_results({"SWEEP1": [-1.0, -0.9], "V(1)": [-1.0, -0.9], "I(Vin)": [1e-12, 2e-12]})
    assert "DC Sweep Analysis Results (2 points):" in dc
    assert "V=-1.000 V   V(1)=-1.000 V  I(Vin)=1.000 pA  " in dc


def test_raw_file_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    res = {"TIME": list(np.cumsum(rng.random(50)) * 1e-6), "V(1)": list(rng.normal(size=50)), "I(Vin)": list(rng.normal(size=50) * 1e-3)}
    p = str(tmp_path / "x.raw")
    R.write_raw(p, res, title="rc sweep instance 3")
    hdr, vals = R.read_raw(p)
    assert hdr["No. Variables"] == "3" and hdr["No. Points"] == "50" and hdr["Plotname"] == "Transient Analysis" and hdr["Flags"] == "real"
    assert np.array_equal(vals["time"], res["TIME"]) and np.array_equal(vals["V(1)"], res["V(1)"]) and np.array_equal(vals["I(Vin)"], res["I(Vin)"])
