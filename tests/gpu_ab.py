"""Development driver (not a pytest file): A/B timing of one deck's transient launch under option sets and
code-shape switches.  Usage:
    python tests/gpu_ab.py <deck> <instances> "<opts>|<defines>|<env>" ...
e.g. python tests/gpu_ab.py rlc 1048576 "share_time_grid=0" "share_time_grid=1" "share_time_grid=1|TSB_X_FOO=1;TSB_X_BAR=0"
<opts>: comma-separated tsb_opts fields; <defines>: $TSB_EXTRA_DEFINES for that variant.  Prints ms per launch (best of 3
after a priming run), steps/s, and the largest deviation of the statistics from the first variant's (0 = bit-identical)."""
import os
import sys

import numpy as np
import torch

import parity_util as PU

T = PU.T


def main():
    deck, n = sys.argv[1], int(sys.argv[2])
    variants = sys.argv[3:] or [""]
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    text = T.BUNDLED[deck]
    ckt = T.Circuit.from_netlist(text, ctx)
    ov = PU.draws(deck, ckt, n)
    dev = {k: torch.from_numpy(v).cuda() for k, v in ov.items()}
    card = ckt.analysis_card()
    ref = None
    for spec in variants:
        o, _, rest = spec.partition("|")
        defs, _, envs = rest.partition("|")
        for item in filter(None, envs.split(";")):          # third field: environment variables for this variant
            k, v = item.split("=")
            os.environ[k] = v
        kw = {}
        for item in filter(None, o.split(",")):
            k, v = item.split("=")
            kw[k] = float(v) if k == "grid_dt" else int(v)
        if defs:
            os.environ["TSB_EXTRA_DEFINES"] = defs
        else:
            os.environ.pop("TSB_EXTRA_DEFINES", None)
        b = ckt.batch(n)
        for (d, p), v in dev.items():
            b.set_param(d, p, v)
        opts = T.default_opts(**kw)
        run = lambda: b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS, opts=opts)
        run(); b.sync()
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); run(); e1.record(stream); stream.synchronize()
            ms.append(e0.elapsed_time(e1))
        tot = b.totals()
        s = np.concatenate([b.stats_all().ravel(), b.counters()[:6].ravel().astype(np.float64), b.rows().astype(np.float64)])
        if ref is None:
            ref = s
        same = np.array_equal(s, ref, equal_nan=True)
        dev_max = 0.0 if same else float(np.nanmax(np.abs(s - ref) / (1e-9 * np.abs(ref) + 1e-12)))
        print(f"{deck} n={n} [{spec:60s}] {min(ms):9.3f} ms  steps/s={tot[0] / (min(ms) * 1e-3):.4e}  bit_identical_to_first={same} dev={dev_max:.3g}", flush=True)
        del b


if __name__ == "__main__":
    main()
