"""Development driver (not a pytest file): how much of the nonlinear kernels' time is warp-mates waiting for each
other's Newton loops?  Runs a deck with the instances in drawn order and sorted by a device parameter (so that lanes of
a warp get similar values and similar iteration counts).  Usage: python tests/gpu_order_experiment.py [deck] [instances]"""
import sys

import numpy as np
import torch

import parity_util as PU

T = PU.T


def main():
    deck = sys.argv[1] if len(sys.argv) > 1 else "diode2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ckt = T.Circuit.from_netlist(T.BUNDLED[deck], ctx)
    ov = PU.draws(deck, ckt, n)
    card = ckt.analysis_card()
    keys = sorted(ov)
    orders = {"drawn order": np.arange(n)}
    for k in keys:
        orders[f"sorted by {k}"] = np.argsort(ov[k], kind="stable")
    for label, perm in orders.items():
        b = ckt.batch(n)
        keep = []
        for k in keys:                                         # parameters stay in drawn order; only the processing order changes
            t = torch.from_numpy(ov[k]).cuda(); keep.append(t)
            b.set_param(k[0], k[1], t)
        b.set_order(None if label == "drawn order" else perm)
        ms = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS)
            e1.record(stream)
            stream.synchronize()
            ms.append(e0.elapsed_time(e1))
        tot = b.totals()
        print(f"{deck} n={n} {label:28s} {min(ms[1:]):8.2f} ms  steps/s={tot[0] / (min(ms[1:]) * 1e-3):.3e}  executed solves={tot[4]}", flush=True)
        del b, keep


if __name__ == "__main__":
    main()
