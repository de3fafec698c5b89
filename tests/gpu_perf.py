"""Development driver (not a pytest file): times one deck's transient launch under kernel-build
variants (precision mode, launch bounds, block size, redundant-solve skipping).
Usage: python tests/gpu_perf.py [deck] [instances]"""
import itertools
import os
import sys
import time

import numpy as np
import torch

import parity_util as PU

T = PU.T


def main():
    deck = sys.argv[1] if len(sys.argv) > 1 else "rlc"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
    variants = sys.argv[3:] or None
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    text = T.BUNDLED[deck]
    ckt = T.Circuit.from_netlist(text, ctx)
    ov = PU.draws(deck, ckt, n)
    dev = {k: torch.from_numpy(v).cuda() for k, v in ov.items()}
    card = ckt.analysis_card()
    grid = [(-1, 1, 0, 128, 1), (-1, 1, 0, 128, 0)]        # library defaults: strict=auto, launch bounds=auto; with / without lane refill
    for strict, skip, mb, bs in itertools.product((0, 1), (1, 0), (1, 2, 3, 4, 5, 6), (64, 128, 256)):
        if bs != 128 and (mb not in (1, 4)):
            continue
        if strict and (skip == 0 or bs != 128):
            continue
        grid.append((strict, skip, mb, bs, 1))
    ref_stats = None
    out_mode = T.OUT_WAVE if os.environ.get("TSB_PERF_OUT") == "wave" else T.OUT_STATS
    cap = int(os.environ.get("TSB_PERF_CAP", "320"))
    for strict, skip, mb, bs, refill in grid:
        tag = f"strict={strict} skip={skip} minblk={mb} block={bs} refill={refill}"
        if variants and not any(v in tag for v in variants):
            continue
        b = ckt.batch(n)
        for (d, p), v in dev.items():
            b.set_param(d, p, v)
        opts = T.default_opts(strict_fp=strict, skip_linear_resolve=skip, min_blocks=mb, block_size=bs, lane_refill=refill)
        try:
            t0 = time.time()
            b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=out_mode, cap_rows=cap, opts=opts)
            b.sync()
            t_first = time.time() - t0
            ms = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=out_mode, cap_rows=cap, opts=opts)
                e1.record(stream)
                stream.synchronize()
                ms.append(e0.elapsed_time(e1))
            tot = b.totals()
            if out_mode == T.OUT_STATS:
                s = b.stats_all()
                if ref_stats is None:
                    ref_stats = s
                dev_max = float(np.nanmax(np.abs(s - ref_stats) / (1e-9 * np.abs(ref_stats) + 1e-12)))
            else:
                nb = int(b.rows().sum()) * b.dims()[1] * 8
                dev_max = nb / (min(ms) * 1e-3) / 1e9        # GB/s of waveform written
            print(f"{deck} n={n} {tag:54s} {min(ms):9.2f} ms  steps/s={tot[0] / (min(ms) * 1e-3):.3e}  first={t_first:.2f}s  "
                  f"dev_vs_first={dev_max:.3g}", flush=True)
        except T.TsbError as e:
            print(tag, "ERROR", str(e)[:300])


if __name__ == "__main__":
    main()
