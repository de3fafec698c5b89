"""CPU-side tests of the product's host logic and C-ABI (no GPU compute calls):
the library loads and exports every symbol include/tspice_b200.h declares; the C++ front-end agrees
with the oracle's independent Python front-end on every deck; numbering, stamped pattern and pivot
order are exact; error behaviour mirrors the reference; the code generator is
deterministic; the product never touches oracle/."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import parity_util as PU

T, O, onl = PU.T, PU.O, PU.onl
ROOT = PU.ROOT
DECKS = sorted(T.BUNDLED)


def test_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "tspice_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tsb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 45
    lib = C.CDLL(T.lib_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(T.api.ABI_SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", T.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (tsb_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert b"sm_100a" in T.lib().tsb_version()


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.TsbError):
        T.Context(0)
    ckt = T.Circuit.from_netlist(T.BUNDLED["rc"])         # host-only plan is fine ...
    b = ckt.batch(4)
    with pytest.raises(T.TsbError):                        # ... but analyses run on the GPU only
        b.run_tran(0.0, 3e-3, 1e-5, 1e-5)
    an = T.NewTransient(0.0, 3e-3, 1e-5, 1e-5, False)
    with pytest.raises(T.TsbError):
        an.Setup(ckt)


def test_product_never_uses_the_oracle():
    pkg = os.path.join(ROOT, "toy-spice_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_kcache" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("# oracle", ""), f"{f} mentions the oracle"


@pytest.mark.parametrize("name", DECKS)
def test_front_end_matches_independent_restatement(built, name):
    """C++ netlist front-end (product) vs Python front-end (oracle): same device table, numbering, cards."""
    ckt = T.Circuit.from_netlist(T.BUNDLED[name])
    oc = O.OracleCircuit(T.BUNDLED[name])
    assert ckt.GetNodeMap() == oc.plan.node_map
    assert ckt.GetBranchMap() == oc.plan.branch_map
    dv = ckt.devices()
    assert len(dv) == len(oc.plan.devices)
    for a, b in zip(dv, oc.plan.devices):
        assert (a["kind"], a["name"], a["nodes"], a["branch"], a["p"], a["ip"]) == (b.kind, b.name, list(b.nodes), b.branch, list(b.p), list(b.ip))
    card, nl = ckt.analysis_card(), oc.netlist
    assert card["analysis"] == nl.analysis
    if nl.analysis == onl.AN_TRAN:
        assert (card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"]) == (
            nl.tran["tstart"], nl.tran["tstop"], nl.tran["tstep"], nl.tran["tmax"], nl.tran["uic"])
    if nl.analysis == onl.AN_DC:
        assert (dv[card["dc_src_dev"]]["name"], card["dc_start"], card["dc_stop"], card["dc_inc"]) == (
            nl.dc["source"], nl.dc["start"], nl.dc["stop"], nl.dc["inc"])
    for an in (T.AN_OP, T.AN_TRAN):
        assert ckt.columns(an) == onl.signal_names(oc.plan, an)


@pytest.mark.parametrize("name", [d for d in DECKS if d != "bjt3"])
def test_symbolic_pass_matches_oracle(built, name):
    """Translate numbering and the first-factor pivot order: product (MarkowitzLU on the nominal
    instance) == oracle (Sparse13 inside a full OP run).  Bit-exact contract."""
    st = T.Circuit.from_netlist(T.BUNDLED[name]).structure()
    so = O.OracleCircuit(T.BUNDLED[name]).structure()
    assert st["ext2int"] == so["ext2int"]
    assert st["pivot_row"] == so["pivot_row"] and st["pivot_col"] == so["pivot_col"]


# SURVEY Appendix A, hand-derived: stamped pattern in first-touch order (OP mode) + transient-only extras
PATTERNS = {
    "rr": ([(3, 1), (1, 3), (1, 1), (1, 2), (2, 1), (2, 2)], []),
    "rl": ([(3, 1), (1, 3), (1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (4, 2), (4, 4)], []),
    "rlc": ([(4, 1), (1, 4), (1, 1), (1, 2), (2, 1), (2, 2), (2, 5), (5, 2), (3, 5), (5, 3), (5, 5), (3, 3)], []),
    "bjt1": ([(4, 1), (1, 4), (1, 1), (1, 2), (2, 1), (2, 2), (1, 3), (3, 1), (3, 3), (3, 2), (2, 3)], []),
    "bjt2": ([(5, 1), (1, 5), (6, 2), (2, 6), (2, 2), (2, 3), (3, 2), (3, 3), (4, 4), (4, 3), (3, 4), (1, 1), (1, 4), (4, 1)], []),
    "mosfet1": ([(4, 1), (1, 4), (5, 2), (2, 5), (1, 1), (1, 3), (3, 1), (3, 3), (3, 2)], [(2, 3), (2, 2)]),
    "transformer1": ([(5, 1), (1, 5), (1, 1), (1, 2), (2, 1), (2, 2), (2, 6), (6, 2), (6, 6), (3, 7), (7, 3), (7, 7), (3, 3), (3, 4),
                      (4, 3), (4, 4)], [(6, 7), (7, 6)]),
    "transformer2": ([(7, 1), (1, 7), (1, 1), (1, 2), (2, 1), (2, 2), (2, 8), (8, 2), (8, 8), (3, 9), (9, 3), (9, 9), (3, 3), (3, 4),
                      (4, 3), (4, 4), (5, 10), (10, 5), (10, 10), (5, 5), (5, 6), (6, 5), (6, 6)],
                     [(8, 9), (9, 8), (8, 10), (10, 8), (9, 10), (10, 9)]),
    "transformer3": ([(5, 1), (1, 5), (1, 1), (1, 2), (2, 1), (2, 2), (2, 6), (6, 2), (6, 6), (3, 3), (3, 4), (4, 3), (4, 4), (3, 7),
                      (7, 3), (7, 7)], [(6, 7), (7, 6)]),
    "vpulse": ([(2, 1), (1, 2), (1, 1)], []),
    "idc": ([(1, 1)], []),
}


@pytest.mark.parametrize("name", sorted(PATTERNS))
def test_stamped_pattern_tables(built, name):
    ckt = T.Circuit.from_netlist(T.BUNDLED[name])
    op, extra = PATTERNS[name]
    assert ckt.pattern(0) == op
    assert ckt.pattern(1) == op + extra


def test_from_devices_equals_from_netlist(built):
    """What a Go host would do: hand over the numbered device table instead of netlist text."""
    a = T.Circuit.from_netlist(T.BUNDLED["transformer2"])
    nn, nb = a.size()
    devs = [(d["kind"], d["name"], d["nodes"], d["branch"], d["p"], d["ip"]) for d in a.devices()]
    b = T.Circuit.from_devices(nn, nb, devs)
    assert a.structure() == b.structure() and a.pattern(1) == b.pattern(1)
    assert a.batch(2).kernel_source() == b.batch(2).kernel_source()


def test_error_behaviour_mirrors_reference(built):
    with pytest.raises(T.TsbError, match="invalid value format"):
        T.Circuit.from_netlist("t\nR1 1 0 abc\n.op\n")
    with pytest.raises(T.TsbError, match="unsupported analysis type"):
        T.Circuit.from_netlist("t\nR1 1 0 1k\n.noise\n")
    with pytest.raises(T.TsbError, match="model not specified|insufficient MOSFET"):
        T.Circuit.from_netlist("t\nM1 1 2 0 0\n.op\n")
    with pytest.raises(T.TsbError, match="not found"):
        T.Circuit.from_netlist("t\nV1 1 0 DC 1\nR1 1 0 1k\n.dc Vx 0 1 0.1\n")
    with pytest.raises(T.TsbError, match="requires exactly 2 nodes|wrong number of nodes"):     # diode.go:45-47 panics
        T.Circuit.from_devices(2, 0, [(T.api.K_D, "d1", [1, 2, 0], 0, [1e-14, 1.0, 0.0], [])])
    with pytest.raises(ValueError, match="inconsistent parameter lengths"):                      # dc.go:21-23 panics
        T.NewDCSweep(["V1"], [0.0, 1.0], [1.0], [0.1])
    with pytest.raises(T.TsbError, match="singular"):
        T.Circuit.from_netlist("t\nV1 1 0 DC 1\nV2 1 0 DC 2\nR1 1 0 1k\n.op\n")   # two sources across one node pair


def test_codegen_is_deterministic_and_specialised(built):
    c1 = T.Circuit.from_netlist(T.BUNDLED["rlc"])
    c2 = T.Circuit.from_netlist(T.BUNDLED["rlc"])
    b1, b2 = c1.batch(8), c2.batch(16)
    assert b1.kernel_source() == b2.kernel_source() and b1.kernel_key() == b2.kernel_key()
    b2.set_param("R1", 0, np.ones(16))
    assert b2.kernel_key() != b1.kernel_key()                     # which parameters vary is compiled in
    assert b1.kernel_key(T.default_opts(strict_fp=1)) != b1.kernel_key()
    src = b2.kernel_source()
    assert "a.pv[0] + inst" in src and "tsb_run_optran_instance<Ckt>" in src
    assert "static constexpr int N = 5;" in src


def test_generated_kernel_compiles_for_sm100a(built, tmp_path):
    """nvcc cross-compiles the specialised unit without a GPU: no local memory, no spills."""
    ckt = T.Circuit.from_netlist(T.BUNDLED["transformer1"])
    b = ckt.batch(2)
    for (dev, par), v in PU.draws("transformer1", ckt, 2).items():
        b.set_param(dev, par, v)
    cu = tmp_path / "k.cu"
    cu.write_text(b.kernel_source())
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-lineinfo",
                        "-Xptxas", "-v", "-cubin", "-o", str(tmp_path / "k.cubin"), str(cu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "0 bytes spill stores, 0 bytes spill loads" in r.stderr
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", str(tmp_path / "k.cubin")], capture_output=True, text=True).stdout
    assert "DFMA" in sass and "sm_100" in sass


def test_destroy_order_is_free(built):
    """Plans keep their context alive and batches their plan (reference counts inside the library): a host whose
    garbage collector releases handles in arbitrary order must not crash.  Host-only here (no context); the GPU
    variant is in test_gpu_parity.py."""
    import ctypes as C
    L = T.lib()
    plan = C.c_void_p()
    assert L.tsb_plan_from_netlist(None, T.BUNDLED["rlc"].encode(), C.byref(plan)) == 0
    batches = []
    for _ in range(3):
        b = C.c_void_p()
        assert L.tsb_batch_create(plan, 16, C.byref(b)) == 0
        batches.append(b)
    L.tsb_plan_destroy(plan)                       # the batches still refer to it
    buf = C.create_string_buffer(64)
    assert L.tsb_batch_kernel_key(batches[0], None, buf, 64) == 0 and len(buf.value) == 32      # plan data still readable
    for b in batches:
        L.tsb_batch_destroy(b)
