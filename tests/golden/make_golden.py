"""Generates tests/golden/golden.npz — regression vectors of the CPU oracle for every bundled deck.

Provenance (read this before trusting them): the reference is Go and cannot be built or run in
this image, and it ships no tests or expected outputs.  These vectors are therefore produced by
oracle/ (the C++ restatement of the reference) — they pin the ORACLE against accidental change and
give the GPU tests fixed numbers to hit; they are NOT outputs of the Go reference ("parity
unpinned").  The independent anchors are in tests/test_oracle.py (analytic rr
answer, hand-derived structure tables, a NumPy restatement of the rc/rl recurrences).

Per deck: 4 instances (instance 0 = nominal values, 1..3 = SURVEY §8(d) draws), a decimated set of
stored rows (<= 48 per instance, always including the first and the last), row counts, status and
step/solve counters.   Run: python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
T = importlib.import_module("toy-spice_b200")
W = importlib.import_module("toy-spice_b200.workloads")
from oracle import oracle as O  # noqa: E402

N_INST = 4
MAX_ROWS = 48


def golden_overrides(name, devices):
    ov = W.sweep_draws(devices, N_INST, W.sweep_seed(name))
    for (dev, par), v in ov.items():          # instance 0 keeps the netlist's nominal value
        nominal = next(d for d in devices if d["name"] == dev)["p"][par]
        v[0] = nominal
    return ov


def pick_rows(nr):
    if nr <= MAX_ROWS:
        return np.arange(nr)
    idx = np.unique(np.concatenate([np.linspace(0, nr - 1, MAX_ROWS).round().astype(int), [0, nr - 1]]))
    return idx


def main():
    out = {}
    for name in sorted(T.BUNDLED):
        if name == "bjt3":       # .ac card: AC analysis is out of scope
            continue
        ckt = T.Circuit.from_netlist(T.BUNDLED[name])
        ov = golden_overrides(name, ckt.devices())
        oc = O.OracleCircuit(T.BUNDLED[name])
        res = oc.run(N_INST, overrides=ov, threads=1, cap_rows=12288)
        ncol = res["ncol"]
        for i in range(N_INST):
            nr = int(res["n_rows"][i])
            idx = pick_rows(nr)
            out[f"{name}/rows_idx/{i}"] = idx.astype(np.int64)
            out[f"{name}/wave/{i}"] = res["wave"][i, idx, :ncol]
        out[f"{name}/n_rows"] = res["n_rows"]
        out[f"{name}/status"] = res["status"]
        out[f"{name}/counters"] = res["counters"][:, :5]
        keys = list(ov.keys())
        out[f"{name}/ov_keys"] = np.array([f"{d}|{p}" for d, p in keys])
        out[f"{name}/ov_vals"] = np.stack([ov[k] for k in keys]) if keys else np.zeros((0, N_INST))
        out[f"{name}/columns"] = np.array(res["signals"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
