"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.

oracle.py — ctypes driver for libtspice_oracle.so (the C++ restatement of the reference solver,
see engine.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  PARITY UNPINNED (no Go toolchain, no reference
golden vectors): the pins this oracle does have are listed in DESIGN.md §Oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

from . import netlist as nlmod

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcJob(C.Structure):
    _fields_ = [("analysis", C.c_int), ("tstart", C.c_double), ("tstop", C.c_double), ("tstep", C.c_double),
                ("tmax", C.c_double), ("uic", C.c_int), ("dc_src_dev", C.c_int), ("dc_start", C.c_double),
                ("dc_stop", C.c_double), ("dc_inc", C.c_double), ("dc2_src_dev", C.c_int), ("dc2_start", C.c_double),
                ("dc2_stop", C.c_double), ("dc2_inc", C.c_double), ("ac_sweep", C.c_int), ("ac_points", C.c_int),
                ("ac_fstart", C.c_double), ("ac_fstop", C.c_double), ("ac_refread", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libtspice_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("tspice_oracle.cpp", "engine.hpp", "sparse13.hpp", "gomath.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libtspice_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_circuit_new.restype = C.c_void_p
        L.orc_circuit_new.argtypes = [C.c_int, C.c_int]
        L.orc_circuit_free.argtypes = [C.c_void_p]
        L.orc_circuit_add.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                      C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int), C.c_int]
        L.orc_n_columns.argtypes = [C.c_void_p, C.c_int]
        L.orc_structure.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_run.argtypes = [C.c_void_p, C.POINTER(OrcJob), C.c_int64, C.c_int, C.POINTER(C.c_int),
                              C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_int64, C.POINTER(C.c_double),
                              C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.orc_go_sin.restype = C.c_double
        L.orc_set_order_sig_buffer.argtypes = [C.c_void_p]
        L.orc_set_order_sig_buffer.restype = None
        L.orc_go_sin.argtypes = [C.c_double]
        L.orc_format_value_factor.argtypes = [C.c_double, C.c_char_p, C.c_int]
        for fn in ("orc_go_max", "orc_go_min", "orc_go_pow"):
            getattr(L, fn).restype = C.c_double
            getattr(L, fn).argtypes = [C.c_double, C.c_double]
        L.orc_lu_batch.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def _arr(ctype, seq):
    return (ctype * max(1, len(seq)))(*seq)


class OracleCircuit:
    """A parsed netlist handed to the CPU oracle.  Mirrors the reference's call sequence
    netlist.Parse -> AssignNodeBranchMaps -> CreateMatrix -> SetupDevices (cmd/spice/main.go:364-401)."""

    def __init__(self, text: str):
        self.netlist = nlmod.parse(text)
        self.plan = nlmod.build_plan(self.netlist)
        L = lib()
        self.h = L.orc_circuit_new(self.plan.n_nodes, self.plan.n_branches)
        for r in self.plan.devices:
            L.orc_circuit_add(self.h, r.kind, r.name.encode(), _arr(C.c_int, r.nodes), len(r.nodes), r.branch,
                              _arr(C.c_double, r.p), len(r.p), _arr(C.c_int, r.ip), len(r.ip))

    def __del__(self):
        try:
            lib().orc_circuit_free(self.h)
        except Exception:
            pass

    @property
    def n(self):
        return self.plan.n_nodes + self.plan.n_branches

    def dev_index(self, name: str) -> int:
        for i, r in enumerate(self.plan.devices):
            if r.name == name:
                return i
        raise KeyError(name)

    def signals(self, analysis=None):
        return nlmod.signal_names(self.plan, self.netlist.analysis if analysis is None else analysis)

    def structure(self):
        n = self.n
        e2i = (C.c_int * (n + 1))(); pr = (C.c_int * (n + 1))(); pc = (C.c_int * (n + 1))()
        rc = lib().orc_structure(self.h, e2i, pr, pc)
        return dict(rc=rc, ext2int=list(e2i)[1:], pivot_row=list(pr)[1:], pivot_col=list(pc)[1:])

    def run(self, n_inst=1, overrides=None, analysis=None, tran=None, dc=None, threads=1, cap_rows=None,
            want_wave=True, want_stats=False, dc2=None, ac=None, ac_refread=False, want_order_sig=False):
        """overrides: {(device_name_or_index, param_index): array[n_inst]}.
        Returns dict(wave [n_inst, cap, ncol], n_rows, status, counters, stats, signals)."""
        nl = self.netlist
        an = nl.analysis if analysis is None else analysis
        job = OrcJob()
        job.analysis = an
        job.dc2_src_dev = -1
        if an == nlmod.AN_TRAN:
            t = dict(nl.tran)
            if tran:
                t.update(tran)
            job.tstart, job.tstop, job.tstep, job.tmax, job.uic = t["tstart"], t["tstop"], t["tstep"], t["tmax"], int(t["uic"])
        elif an == nlmod.AN_DC:
            d = dict(nl.dc)
            if dc:
                d.update(dc)
            job.dc_src_dev = self.dev_index(d["source"])
            job.dc_start, job.dc_stop, job.dc_inc = d["start"], d["stop"], d["inc"]
            if dc2:         # nested sweep (dc.go:205-270): dict(source, start, stop, inc) of the INNER source
                job.dc2_src_dev = self.dev_index(dc2["source"])
                job.dc2_start, job.dc2_stop, job.dc2_inc = dc2["start"], dc2["stop"], dc2["inc"]
        elif an == nlmod.AN_AC:
            a = dict(nl.ac)
            if ac:
                a.update(ac)
            job.ac_sweep = {"DEC": 0, "OCT": 1, "LIN": 2}[a["sweep"].upper()]
            job.ac_points, job.ac_fstart, job.ac_fstop = int(a["points"]), a["fstart"], a["fstop"]
            job.ac_refread = int(bool(ac_refread))
        overrides = overrides or {}
        keys = list(overrides.keys())
        ov_dev = [self.dev_index(k[0]) if isinstance(k[0], str) else int(k[0]) for k in keys]
        ov_par = [int(k[1]) for k in keys]
        vals = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(overrides[k], dtype=np.float64), (n_inst,))
                                              for k in keys]) if keys else np.zeros((0, n_inst)))
        nested = an == nlmod.AN_DC and job.dc2_src_dev >= 0
        ncol = lib().orc_n_columns(self.h, 4 if nested else an)
        if cap_rows is None:
            cap_rows = 1 if an == nlmod.AN_OP else (int(round((job.dc_stop - job.dc_start) / job.dc_inc)) + 3
                                                     if an == nlmod.AN_DC else job.ac_points if an == nlmod.AN_AC else 65536)
            if nested:
                cap_rows *= int(round((job.dc2_stop - job.dc2_start) / job.dc2_inc)) + 3
        wave = np.full((n_inst, cap_rows, ncol), np.nan) if want_wave else None
        stats = np.zeros((n_inst, 4, ncol)) if want_stats else None
        n_rows = np.zeros(n_inst, dtype=np.int64)
        status = np.zeros(n_inst, dtype=np.int32)
        counters = np.zeros((n_inst, 6), dtype=np.int64)
        dp = C.POINTER(C.c_double)
        order_sig = np.zeros(n_inst, dtype=np.uint64) if want_order_sig else None
        if want_order_sig:       # per instance: signature of the pivot choices the reference makes on ITS values (sparse13.hpp)
            lib().orc_set_order_sig_buffer(order_sig.ctypes.data_as(C.c_void_p))
        rc = lib().orc_run(self.h, C.byref(job), n_inst, len(keys), _arr(C.c_int, ov_dev), _arr(C.c_int, ov_par),
                           vals.ctypes.data_as(dp), threads, cap_rows,
                           wave.ctypes.data_as(dp) if wave is not None else None,
                           n_rows.ctypes.data_as(C.POINTER(C.c_int64)), status.ctypes.data_as(C.POINTER(C.c_int32)),
                           counters.ctypes.data_as(C.POINTER(C.c_int64)),
                           stats.ctypes.data_as(dp) if stats is not None else None)
        if want_order_sig:
            lib().orc_set_order_sig_buffer(None)
        if rc != 0:
            raise RuntimeError(f"orc_run failed rc={rc}")
        return dict(wave=wave, n_rows=n_rows, status=status, counters=counters, stats=stats, order_sig=order_sig,
                    signals=(["SWEEP1", "SWEEP2"] + self.signals(an)[1:]) if nested else self.signals(an), ncol=ncol)


def go_sin(x: float) -> float:
    return lib().orc_go_sin(x)


def format_value_factor(v: float) -> str:
    buf = C.create_string_buffer(64)
    lib().orc_format_value_factor(v, buf, 64)
    return buf.value.decode()


def lu_batch(A_nominal: np.ndarray, A: np.ndarray, b: np.ndarray):
    """Sparse 1.3 restatement on a batch of dense systems, order frozen on `A_nominal` (orc_lu_batch).
    Returns x [n_inst, n], status [n_inst], (pivot_row, pivot_col) 1-based."""
    n = A_nominal.shape[0]
    A = np.ascontiguousarray(A, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    An = np.ascontiguousarray(A_nominal, dtype=np.float64)
    n_inst = A.shape[0]
    x = np.zeros((n_inst, n)); st = np.zeros(n_inst, dtype=np.int32)
    pr = (C.c_int * n)(); pc = (C.c_int * n)()
    dp = C.POINTER(C.c_double)
    rc = lib().orc_lu_batch(n, An.ctypes.data_as(dp), n_inst, A.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp),
                            st.ctypes.data_as(C.POINTER(C.c_int32)), pr, pc)
    if rc != 0:
        raise RuntimeError("oracle: nominal matrix is singular")
    return x, st, (np.array(pr[:]), np.array(pc[:]))
