"""Development driver (not a pytest file): every bundled deck through the GPU path and the CPU oracle
on the same draws; prints one parity line per deck.  Usage: python tests/gpu_sweep.py [n] [deck ...]"""
import json
import sys
import tempfile
import time

import parity_util as PU

T = PU.T


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    decks = sys.argv[2:] or [d for d in sorted(T.BUNDLED) if d != "bjt3"]
    ctx = T.Context(0)
    print("SMs", ctx.sm_count, "fp64 peak TF", round(ctx.measure_fp64_peak(), 2))
    bad = 0
    for strict in (0, 1):
        for name in decks:
            text = T.BUNDLED[name]
            ckt0 = T.Circuit.from_netlist(text)
            ov = PU.draws(name, ckt0, n)
            opts = T.default_opts(strict_fp=strict)
            t0 = time.time()
            try:
                ckt, batch, an = PU.run_gpu(ctx, text, n, ov, cap_rows=12288, opts=opts)
            except T.TsbError as e:
                print(name, "GPU ERROR", e)
                bad += 1
                continue
            t1 = time.time()
            _, ores = PU.run_oracle(text, n, ov, cap_rows=12288)
            t2 = time.time()
            rep = PU.compare_waves(batch, ores, n)
            ok = PU.report_ok(rep)
            bad += 0 if ok else 1
            tot = batch.totals()
            print(f"{'strict' if strict else 'fma   '} {name:13s} {'OK ' if ok else 'BAD'} gpu {t1 - t0:6.2f}s cpu {t2 - t1:6.2f}s totals={tot.tolist()} "
                  f"{json.dumps({k: v for k, v in rep.items() if k != 'n'})}", flush=True)
    # NVRTC path: empty cache directory forces a run-time compile
    with tempfile.TemporaryDirectory() as d:
        ctx2 = T.Context(0)
        ctx2.set_cache_dir(d)
        text = T.BUNDLED["rlc"]
        ckt0 = T.Circuit.from_netlist(text)
        ov = PU.draws("rlc", ckt0, 32)
        t0 = time.time()
        ckt, batch, an = PU.run_gpu(ctx2, text, 32, ov, cap_rows=12288)
        _, ores = PU.run_oracle(text, 32, ov, cap_rows=12288)
        rep = PU.compare_waves(batch, ores, 32)
        print("NVRTC path rlc", "OK" if PU.report_ok(rep) else "BAD", round(time.time() - t0, 2), "s", rep)
        bad += 0 if PU.report_ok(rep) else 1
    print("FAILED DECKS:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
