"""Development driver (not a pytest file): transient throughput of RC ladders of growing order — the thread-per-circuit
kernels against the cooperative mapping (tsb_opts.coop_parts = 2 / 4: one instance per 2 / 4 threads in different warps).
Usage: python tests/gpu_ladder_perf.py [instances] [sections,...] [TSB_COOP_MIN_BLOCKS values, comma-separated]"""
import os
import sys

import numpy as np
import torch

import parity_util as PU
import random_decks as _RD
rc_ladder = getattr(_RD, os.environ.get("LADDER_KIND", "rc_ladder"))       # e.g. LADDER_KIND=diode_rc_ladder

T = PU.T


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    secs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 8, 12, 16, 24]
    mbs = [x for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [""]
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    for sections in secs:
        text = rc_ladder(sections)
        ckt = T.Circuit.from_netlist(text, ctx)
        ov = PU.draws("ladder", ckt, n, seed=5)
        dev = {k: torch.from_numpy(v).cuda() for k, v in ov.items()}
        card = ckt.analysis_card()
        ref = None
        for parts in (0, 2, 4, 8):
            if parts and ckt.coop_info(parts) is None:
                continue
            for mb in (mbs if parts else [""]):
                if mb:
                    os.environ["TSB_EXTRA_DEFINES"] = f"TSB_COOP_MIN_BLOCKS={mb}"
                else:
                    os.environ.pop("TSB_EXTRA_DEFINES", None)
                b = ckt.batch(n)
                for (d, p), v in dev.items():
                    b.set_param(d, p, v)
                o = T.default_opts(coop_parts=parts)
                ms = []
                try:
                    for it in range(3):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS, opts=o)
                        e1.record(stream)
                        stream.synchronize()
                        ms.append(e0.elapsed_time(e1))
                except T.TsbError as ex:
                    print(f"ladder sections={sections:2d} n={ckt.n:2d} instances={n} coop_parts={parts}: does not launch ({str(ex)[:120]})", flush=True)
                    del b
                    continue
                tot = b.totals()
                s = b.stats_all()
                if ref is None:
                    ref = s
                dev_max = float(np.nanmax(np.abs(s[[0, 1, 3]] - ref[[0, 1, 3]]) / (1e-9 * np.abs(ref[[0, 1, 3]]) + 1e-12)))
                t = min(ms[1:]) * 1e-3
                print(f"ladder sections={sections:2d} n={ckt.n:2d} instances={n} coop_parts={parts} min_blocks={mb or 'default':7s} {min(ms[1:]):9.2f} ms  "
                      f"steps/s={tot[0] / t:.3e}  failed={int((b.status() != 0).sum())}  dev_vs_thread_mapping={dev_max:.3g} (1 = contract)", flush=True)
                del b


if __name__ == "__main__":
    main()
