"""Result formats on top of GetResults(): the reference CLI's table output and a SPICE raw file.

  format_value_factor   pkg/util/formatter.go:8-24   ("%.3f <prefix><unit>", "%.3e" below 1e-12)
  format_results        cmd/spice/main.go:17-185     printResults for operating point / DC sweep / transient results
                        (same headers, same column order: V(...) then I(...), each sorted by name; AC is out of scope)
  write_raw             ASCII "rawfile" as SPICE3 / ngspice write it (not in the reference; for waveform viewers)

Host-side presentation only: no computation happens here."""
from __future__ import annotations

import datetime as _dt
from typing import Mapping, Sequence


def _f3(x: float) -> str:
    """fmt.Sprintf("%.3f", x) — Go and Python round the exact binary value the same way (half to even on ties)."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "+Inf" if x > 0 else "-Inf"
    return f"{x:.3f}"


def format_value_factor(value: float, unit: str) -> str:
    a = abs(value)
    if value != value:                       # NaN: every comparison is false -> the default branch
        return f"NaN {unit}"
    if a >= 1:
        return f"{_f3(value)} {unit}"
    if a >= 1e-3:
        return f"{_f3(value * 1e3)} m{unit}"
    if a >= 1e-6:
        return f"{_f3(value * 1e6)} u{unit}"
    if a >= 1e-9:
        return f"{_f3(value * 1e9)} n{unit}"
    if a >= 1e-12:
        return f"{_f3(value * 1e12)} p{unit}"
    return f"{value:.3e} {unit}"


def _names(results: Mapping[str, Sequence[float]], skip=()):
    v = sorted(n for n in results if n.startswith("V(") and n not in skip)
    i = sorted(n for n in results if n.startswith("I(") and n not in skip)
    return v, i


def format_results(results: Mapping[str, Sequence[float]]) -> str:
    """The text the reference's `spice <netlist>` prints for this GetResults() map."""
    out = ["", "Analysis Results:", "================"]
    if "SWEEP1" in results:
        sweep = results["SWEEP1"]
        out += ["", f"DC Sweep Analysis Results ({len(sweep)} points):", "Sweep Values    Node Voltages        Branch Currents",
                "------------------------------------------------"]
        vn, cn = _names(results, ("SWEEP1", "SWEEP2"))
        for k in range(len(sweep)):
            line = "V=%-9s  " % format_value_factor(sweep[k], "V")
            line += "".join(f"{n}={format_value_factor(results[n][k], 'V')}  " for n in vn)
            line += "".join(f"{n}={format_value_factor(results[n][k], 'A')}  " for n in cn)
            out.append(line)
        return "\n".join(out) + "\n"
    if len(results.get("TIME", ())) <= 1:
        vn, cn = _names(results)
        out += ["", "Node Voltages:"] + [f"{n} = {format_value_factor(results[n][0], 'V')}" for n in vn]
        out += ["", "Branch Currents:"] + [f"{n} = {format_value_factor(results[n][0], 'A')}" for n in cn]
        return "\n".join(out) + "\n"
    times = results["TIME"]
    out += ["", f"Transient Analysis Results ({len(times)} time points):", "Time        Node Voltages        Branch Currents",
            "------------------------------------------------"]
    vn, cn = _names(results, ("TIME",))
    for k, t in enumerate(times):
        line = "%9s  " % format_value_factor(t, "s")
        line += "".join(f"{n}={format_value_factor(results[n][k], 'V')}  " for n in vn)
        line += "".join(f"{n}={format_value_factor(results[n][k], 'A')}  " for n in cn)
        out.append(line)
    return "\n".join(out) + "\n"


def write_raw(path: str, results: Mapping[str, Sequence[float]], title: str = "tspice_b200", columns: Sequence[str] | None = None):
    """ASCII rawfile (Title / Date / Plotname / Flags / No. Variables / No. Points / Variables / Values).  `columns`
    fixes the variable order (default: the axis TIME / SWEEP1 first, then the map's own order)."""
    axis = "TIME" if "TIME" in results else ("SWEEP1" if "SWEEP1" in results else None)
    cols = list(columns) if columns else ([axis] if axis else []) + [n for n in results if n != axis]
    npts = len(results[cols[0]]) if cols else 0
    plot = "Transient Analysis" if axis == "TIME" and npts > 1 else ("DC transfer characteristic" if axis == "SWEEP1" else "Operating Point")

    def kind(n):
        if n == "TIME":
            return "time"
        return "current" if n.startswith("I(") else "voltage"

    with open(path, "w") as f:
        f.write(f"Title: {title}\nDate: {_dt.datetime.now().ctime()}\nPlotname: {plot}\nFlags: real\n")
        f.write(f"No. Variables: {len(cols)}\nNo. Points: {npts}\nVariables:\n")
        for k, n in enumerate(cols):
            f.write(f"\t{k}\t{n.lower() if n in ('TIME', 'SWEEP1') else n}\t{kind(n)}\n")
        f.write("Values:\n")
        for p in range(npts):
            f.write(f"{p}\t{results[cols[0]][p]:.16e}\n")
            for n in cols[1:]:
                f.write(f"\t{results[n][p]:.16e}\n")
            f.write("\n")


def read_raw(path: str):
    """Reader for write_raw's output (round-trip tests): returns (header dict, {variable: [values]})."""
    hdr, names, vals = {}, [], {}
    lines = open(path).read().split("\n")
    i = 0
    while i < len(lines) and lines[i] != "Variables:":
        if ":" in lines[i]:
            k, v = lines[i].split(":", 1)
            hdr[k.strip()] = v.strip()
        i += 1
    i += 1
    while lines[i] != "Values:":
        names.append(lines[i].split("\t")[2])
        i += 1
    i += 1
    vals = {n: [] for n in names}
    tok = [t for ln in lines[i:] for t in ln.split("\t") if t.strip()]
    nv = len(names)
    p = 0
    while p + nv < len(tok) + 1 and p < len(tok):
        row = tok[p + 1: p + 1 + nv]
        for n, t in zip(names, row):
            vals[n].append(float(t))
        p += 1 + nv
    return hdr, vals
