"""Shared helpers for the parity tests: run one workload deck through the product (GPU, C-ABI) and
through the CPU oracle on the same seeded parameter draws, and compare.

Tolerance contract (BASELINE.json north_star / SURVEY.md §8c): node numbering and stamped pattern
exact; voltages and currents |gpu - ref| <= 1e-9*|ref| + 1e-12 at identical time points."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T = importlib.import_module("toy-spice_b200")
from oracle import netlist as onl  # noqa: E402
from oracle import oracle as O  # noqa: E402

RELTOL, ABSTOL = 1e-9, 1e-12


def log_uniform(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))


def draws(name: str, ckt, n: int, seed: int | None = None):
    """Parameter draws of SURVEY.md §8(d) (toy-spice_b200/workloads.py)."""
    from importlib import import_module
    W = import_module("toy-spice_b200.workloads")
    return W.sweep_draws(ckt.devices(), n, W.sweep_seed(name) if seed is None else seed)


def run_gpu(ctx, text, n, overrides, out=T.OUT_WAVE, cap_rows=0, opts=None, analysis=None, tran=None, grid_dt=0.0, dc=None, dc2=None):
    """Through the reference-shaped API: Circuit -> analysis.Setup(batch) -> Execute.
    dc = (source, start, stop, inc) overrides the deck's .dc card; dc2 = (outer, inner) tuples runs the nested sweep."""
    ckt = T.Circuit.from_netlist(text, ctx)
    batch = ckt.batch(n)
    for (dev, par), vals in overrides.items():
        batch.set_param(dev, par, vals)
    card = ckt.analysis_card()
    if tran:
        card.update(tran)
    an_kind = card["analysis"] if analysis is None else analysis
    if an_kind == T.AN_OP:
        an = T.NewOP()
    elif an_kind == T.AN_TRAN:
        an = T.NewTransient(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"])
        an.out = out
        an.cap_rows = cap_rows
        an.grid_dt = grid_dt
    else:
        if dc2 is not None:
            (s1, a1, b1, c1), (s2, a2, b2, c2) = dc2
            an = T.NewDCSweep([s1, s2], [a1, a2], [b1, b2], [c1, c2])
        elif dc is not None:
            an = T.NewDCSweep([dc[0]], [dc[1]], [dc[2]], [dc[3]])
        else:
            an = T.NewDCSweep([ckt.devices()[card["dc_src_dev"]]["name"]], [card["dc_start"]], [card["dc_stop"]], [card["dc_inc"]])
        an.out = out
    if opts is not None:
        an.opts = opts
    an.Setup(batch)
    an.Execute()
    return ckt, batch, an


def run_oracle(text, n, overrides, threads=0, cap_rows=None, want_stats=False, analysis=None, tran=None, want_wave=True, dc=None,
               dc2=None, **extra):
    oc = O.OracleCircuit(text)
    kw = {}
    if dc2 is not None:
        (s1, a1, b1, c1), (s2, a2, b2, c2) = dc2
        kw = dict(dc=dict(source=s1, start=a1, stop=b1, inc=c1), dc2=dict(source=s2, start=a2, stop=b2, inc=c2))
    elif dc is not None:
        kw = dict(dc=dict(source=dc[0], start=dc[1], stop=dc[2], inc=dc[3]))
    res = oc.run(n, overrides=overrides, threads=threads, cap_rows=cap_rows, want_stats=want_stats, analysis=analysis, tran=tran,
                 want_wave=want_wave, **kw, **extra)
    return oc, res


def compare_waves(batch, ores, n, label="", reltol=RELTOL, abstol=ABSTOL, nonfinite_any=False):
    """Instance-by-instance comparison at identical stored rows.  Returns a report dict
    (max_rel = worst |gpu - ref| / (reltol*|ref| + abstol); <= 1 means inside the tolerance)."""
    rows_g = batch.rows()
    st_g = batch.status()
    cnt_g = batch.counters()
    ncol = ores["ncol"]
    rep = dict(n=n, row_mismatch=0, status_mismatch=0, max_rel=0.0, max_abs=0.0, worst=None, nan_mismatch=0, inf_mismatch=0,
               compared_points=0, counter_mismatch=0, counter_flips=[], counter_mismatch_failed_lanes=0)
    wall = batch.wave_all() if n > 64 else None
    for i in range(n):
        nr_o = int(ores["n_rows"][i])
        if int(st_g[i]) != int(ores["status"][i]):
            rep["status_mismatch"] += 1
            continue
        if int(rows_g[i]) != nr_o:
            rep["row_mismatch"] += 1
            continue
        if not np.array_equal(cnt_g[:4, i], ores["counters"][i, :4]):
            if int(st_g[i]) != 0:
                # a lane whose Newton iteration does NOT converge (status != 0 on both sides) runs a chaotic fixed-point
                # map for up to 100 iterations per attempt: last-bit differences (libm exp / pow, FMA) change how many
                # of the fallback stages converge on the way to the same failure.  Counted apart, not held to equality.
                rep["counter_mismatch_failed_lanes"] += 1
            else:
                rep["counter_mismatch"] += 1
                rep["counter_flips"].append((i, cnt_g[:4, i].tolist(), ores["counters"][i, :4].tolist()))
        wg = wall[:nr_o, :, i] if wall is not None else batch.waveform(i)
        wo = ores["wave"][i, :nr_o, :ncol]
        nan_g, nan_o = np.isnan(wg), np.isnan(wo)
        if nonfinite_any:
            # a deck that overflows on purpose: whether an overflowed lane reads Inf or NaN (Inf - Inf, 0 * Inf) depends on the
            # last bit of exp() and on the elimination order; only finite / non-finite is compared
            if not np.array_equal(np.isfinite(wg), np.isfinite(wo)):
                rep["nan_mismatch"] += 1
                continue
            ok = np.isfinite(wo)
            err = np.abs(wg[ok] - wo[ok]); tol = reltol * np.abs(wo[ok]) + abstol
            rep["compared_points"] += int(ok.sum())
            if err.size and float((err / tol).max()) > rep["max_rel"]:
                rep["max_rel"] = float((err / tol).max()); rep["max_abs"] = float(err.max())
            continue
        if not np.array_equal(nan_g, nan_o):
            rep["nan_mismatch"] += 1
            r, c = np.argwhere(nan_g != nan_o)[0]
            rep.setdefault("first_class_mismatch", ("nan", i, int(r), int(c), float(wg[r, c]), float(wo[r, c])))
            continue
        # infinities are results too (BJT overflow lanes): same places, same signs
        inf_g, inf_o = np.isinf(wg), np.isinf(wo)
        if not np.array_equal(inf_g, inf_o) or not np.array_equal(np.sign(wg[inf_g]), np.sign(wo[inf_o])):
            rep["inf_mismatch"] += 1
            r, c = np.argwhere((inf_g != inf_o) | (inf_g & inf_o & (np.sign(wg) != np.sign(wo))))[0]
            rep.setdefault("first_class_mismatch", ("inf", i, int(r), int(c), float(wg[r, c]), float(wo[r, c])))
            continue
        ok = ~nan_o & np.isfinite(wo) & np.isfinite(wg)
        err = np.abs(wg[ok] - wo[ok])
        tol = reltol * np.abs(wo[ok]) + abstol
        rep["compared_points"] += int(ok.sum())
        if err.size:
            ratio = err / tol
            j = int(np.argmax(ratio))
            if ratio[j] > rep["max_rel"]:
                rep["max_rel"] = float(ratio[j])
                rep["max_abs"] = float(err[j])
                rep["worst"] = (i, float(wg[ok][j]), float(wo[ok][j]))
    return rep


def report_str(rep) -> str:
    """The whole report on one line (pytest abbreviates dict reprs in assertion messages)."""
    return "; ".join(f"{k}={v}" for k, v in rep.items() if k != "counter_flips" or v)


def report_ok(rep) -> bool:
    return (rep["row_mismatch"] == 0 and rep["status_mismatch"] == 0 and rep["nan_mismatch"] == 0 and rep["inf_mismatch"] == 0
            and rep["max_rel"] <= 1.0)


def resample_reference(wave, n_rows, ncol, tg):
    """The TSB_OUT_GRID definition applied (in numpy) to one instance's reference series `wave[:n_rows, :ncol]`:
    linear interpolation between consecutive stored rows, constant before the first / after the last one."""
    t = wave[:n_rows, 0]
    out = np.empty((len(tg), ncol))
    out[:, 0] = tg
    i1 = np.searchsorted(t, tg, side="left")            # first stored row with t >= tg
    for k, g in enumerate(tg):
        j = int(i1[k])
        if j >= n_rows:
            out[k, 1:] = wave[n_rows - 1, 1:ncol]
        elif j == 0:
            out[k, 1:] = wave[0, 1:ncol]
        else:
            w = (g - t[j - 1]) / (t[j] - t[j - 1])
            out[k, 1:] = wave[j - 1, 1:ncol] + (wave[j, 1:ncol] - wave[j - 1, 1:ncol]) * w
    return out
