"""Development driver (not a pytest file): transient throughput of RC ladders of growing order through the
thread-per-circuit analysis kernels — where that mapping stops paying.  Usage: python tests/gpu_ladder_perf.py [instances]"""
import sys

import torch

import parity_util as PU
from random_decks import rc_ladder

T = PU.T


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    for sections in (1, 4, 8, 12, 16, 24):
        text = rc_ladder(sections)
        ckt = T.Circuit.from_netlist(text, ctx)
        ov = PU.draws("ladder", ckt, n, seed=5)
        b = ckt.batch(n)
        for (d, p), v in ov.items():
            b.set_param(d, p, torch.from_numpy(v).cuda())
        card = ckt.analysis_card()
        ms = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS)
            e1.record(stream)
            stream.synchronize()
            ms.append(e0.elapsed_time(e1))
        tot = b.totals()
        t = min(ms[1:]) * 1e-3
        print(f"ladder sections={sections:2d} n={ckt.n:2d} instances={n}  {min(ms[1:]):9.2f} ms  steps/s={tot[0] / t:.3e}  "
              f"solves/s={tot[4] / t:.3e}  failed={int((b.status() != 0).sum())}", flush=True)
        del b


if __name__ == "__main__":
    main()
