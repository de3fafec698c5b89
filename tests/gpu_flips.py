"""Development driver (not a pytest file): the discrete-decision differences between the CUDA path and the oracle, per deck
and build — Newton solve counts (and, if any, accepted / rejected step counts) of every instance of the 48-instance parity
sweeps and of a larger 4096-instance sweep of the nonlinear decks.  The CUDA exp / log / pow differ from glibc's by <= 1 ulp on
some arguments, exactly as Go's do: the flip rate measured here is the best available estimate of how often a count of the
real Go solver would differ from the oracle's.  Usage: python tests/gpu_flips.py > gpurun_out/flips.txt"""
import numpy as np

import parity_util as PU
from extra_decks import EXTRA

T, O = PU.T, PU.O


def main():
    ctx = T.Context(0)
    decks = {n: (T.BUNDLED[n], {}) for n in sorted(T.BUNDLED) if n != "bjt3"}
    tran150 = dict(tstart=0.0, tstop=150e-6, tstep=1e-6, tmax=0.0, uic=False)
    for n in ("bjt1", "bjt3"):
        decks[n + "_tran"] = (T.BUNDLED[n], dict(analysis=T.AN_TRAN, tran=tran150))
    for n, (text, an) in sorted(EXTRA.items()):
        if "tran" in an or text.lower().find(".tran") >= 0:
            decks["x:" + n] = (text, {})
    print("deck                build   n     status!= rows!=  solve-count flips (instances)   of which step-count flips   failed-lane flips")
    for name, (text, kw) in decks.items():
        ckt0 = T.Circuit.from_netlist(text)
        nonlinear = any(d["kind"] in (5, 6, 7) for d in ckt0.devices())
        for n in ((48, 4096) if nonlinear else (48,)):
            ov = PU.draws(name.split(":")[-1].replace("_tran", ""), ckt0, n)
            okw = dict(kw)
            if "analysis" in okw:
                okw["analysis"] = 1
            _, ores = PU.run_oracle(text, n, ov, cap_rows=12288, want_wave=False, **okw)
            for mode, label in ((-1, "auto"), (1, "strict"), (0, "fast")):
                try:
                    _, b, _ = PU.run_gpu(ctx, text, n, ov, out=T.OUT_STATS if kw.get("analysis", ckt0.analysis_card()["analysis"]) == T.AN_TRAN else T.OUT_WAVE,
                                         cap_rows=12288, opts=T.default_opts(strict_fp=mode), **kw)
                except Exception as ex:
                    print(f"{name:18s} {label:6s} {n:5d}  error {str(ex)[:60]}")
                    continue
                cg, co = b.counters()[:4].T, ores["counters"][:, :4]
                stg, sto = b.status(), ores["status"]
                st_bad = int((stg != sto).sum())
                ok = stg == sto
                row_bad = int((b.rows()[ok] != ores["n_rows"][ok]).sum())
                good = ok & (sto == 0)
                flips = np.any(cg != co, axis=1)
                step_flips = np.any(cg[:, :2] != co[:, :2], axis=1)
                print(f"{name:18s} {label:6s} {n:5d}  {st_bad:7d} {row_bad:6d}  {int((flips & good).sum()):6d}"
                      f"                          {int((step_flips & good).sum()):6d}                     {int((flips & ok & (sto != 0)).sum()):6d}", flush=True)
                del b


if __name__ == "__main__":
    main()
