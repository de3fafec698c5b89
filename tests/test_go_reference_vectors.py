"""Pinning hook: outputs of the REAL Go reference (baseline/go/main.go -> tests/golden/go/<deck>.json), when
somebody with a Go toolchain has produced them, are the golden vectors for the oracle and for the CUDA path.
None can be produced in this image (no Go, sparse module not vendored): the tests then skip, and DESIGN.md says
"parity unpinned"."""
import glob
import json
import os

import numpy as np
import pytest

import parity_util as PU

T, O = PU.T, PU.O
FILES = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "go", "*.json")))
SPECIAL = {"NaN": np.nan, "+Inf": np.inf, "-Inf": -np.inf}


def _load(path):
    d = json.load(open(path))
    res = {k: np.array([SPECIAL.get(v, v) if isinstance(v, str) else v for v in col], dtype=np.float64) for k, col in d["results"].items()}
    return d["deck"], d["analysis"], d.get("error", ""), res


def _check(series_by_name, res, label):
    for key, ref in res.items():
        got = series_by_name[key]
        assert len(got) == len(ref), (label, key, len(got), len(ref))
        if key in ("TIME", "SWEEP1"):
            assert np.array_equal(got, ref), (label, key)
            continue
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (label, key)
        ok = np.isfinite(ref)
        assert np.all(np.abs(got[ok] - ref[ok]) <= PU.RELTOL * np.abs(ref[ok]) + PU.ABSTOL), (label, key)


@pytest.mark.skipif(not FILES, reason="no Go-reference vectors in tests/golden/go (no Go toolchain in this image)")
@pytest.mark.parametrize("path", FILES)
def test_oracle_matches_the_go_reference(path):
    deck, kind, err, res = _load(path)
    if deck not in T.BUNDLED:
        pytest.skip(f"{deck}: not a bundled deck")
    oc = O.OracleCircuit(T.BUNDLED[deck])
    r = oc.run(1, cap_rows=24000)
    names = r["signals"]
    nr = int(r["n_rows"][0])
    series = {n: r["wave"][0, :nr, j] for j, n in enumerate(names)}
    assert (int(r["status"][0]) != 0) == bool(err), (deck, err)
    _check(series, res, deck)


@pytest.mark.gpu
@pytest.mark.skipif(not FILES, reason="no Go-reference vectors in tests/golden/go (no Go toolchain in this image)")
@pytest.mark.parametrize("path", FILES)
def test_cuda_path_matches_the_go_reference(ctx, path):
    deck, kind, err, res = _load(path)
    if deck not in T.BUNDLED:
        pytest.skip(f"{deck}: not a bundled deck")
    ckt, batch, an = PU.run_gpu(ctx, T.BUNDLED[deck], 1, {}, cap_rows=24000, opts=T.default_opts(strict_fp=1))
    series = an.GetResults(0)
    _check({k: np.asarray(v) for k, v in series.items()}, res, deck)
