"""Development driver (not a pytest file): ONE transient launch of a deck after a priming run — the target of `ncu`.
Usage: python tests/gpu_one.py <deck> <instances> [opts k=v,k=v] [out=stats|wave|grid]"""
import sys

import torch

import parity_util as PU

T = PU.T


def main():
    deck, n = sys.argv[1], int(sys.argv[2])
    kw = {}
    for item in filter(None, (sys.argv[3] if len(sys.argv) > 3 else "").split(",")):
        k, v = item.split("=")
        kw[k] = int(v)
    out = {"stats": T.OUT_STATS, "wave": T.OUT_WAVE, "grid": T.OUT_GRID}[sys.argv[4] if len(sys.argv) > 4 else "stats"]
    ctx = T.Context(0)
    if deck.startswith("ladder"):
        from random_decks import rc_ladder
        text = rc_ladder(int(deck[6:]))
    elif deck.startswith("diodeladder"):
        from random_decks import diode_rc_ladder
        text = diode_rc_ladder(int(deck[11:]))
    else:
        text = T.BUNDLED[deck]
    ckt = T.Circuit.from_netlist(text, ctx)
    ov = PU.draws(deck, ckt, n, seed=5 if "ladder" in deck else None)
    dev = {k: torch.from_numpy(v).cuda() for k, v in ov.items()}
    torch.cuda.synchronize()
    card = ckt.analysis_card()
    b = ckt.batch(n)
    for (d, p), v in dev.items():
        b.set_param(d, p, v)
    opts = T.default_opts(**kw)
    for _ in range(2):
        b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=out, cap_rows=320, opts=opts)
        b.sync()
    print(deck, n, kw, "totals", b.totals().tolist())


if __name__ == "__main__":
    main()
