"""Development driver (not a pytest file) for `compute-sanitizer --tool memcheck`: one small run of every kernel family
(linear / nonlinear transient with statistics, waveform and fixed-grid output, lane refill, OP, DC sweep, the
operator-level LU at each lane width) so that out-of-bounds accesses and misaligned shared-memory traffic surface."""
import numpy as np

import parity_util as PU
from test_lu_operator import mna_like

T = PU.T


def main():
    ctx = T.Context(0)
    for name in ("rc", "rlc", "transformer2", "diode2", "diode3", "diode1", "bjt2", "mosfet1", "vpwl"):
        text = T.BUNDLED[name]
        n = 200                                             # not a multiple of the block size: tail warps and blocks
        ov = PU.draws(name, T.Circuit.from_netlist(text), n)
        for out, kw in ((T.OUT_STATS, {}), (T.OUT_WAVE, {"cap_rows": 64}), (T.OUT_GRID, {"grid_dt": 0.0})):
            card = T.Circuit.from_netlist(text).analysis_card()
            if card["analysis"] != T.AN_TRAN and out == T.OUT_GRID:
                continue
            for refill in (0, 1):
                if refill and name not in ("diode2", "bjt2"):
                    continue
                _, b, _ = PU.run_gpu(ctx, text, n, ov, out=out, opts=T.default_opts(lane_refill=refill), **kw)
                b.status(); b.rows()
                del b
        print(name, "ok", flush=True)
    for n in (1, 5, 8, 9, 16, 17, 32):
        base, A, b = mna_like(n, 333, n)
        order = T.lu_order(base)
        for strict in (False, True):
            x, st = ctx.lu_solve_batched(A, b, order, strict=strict)
            assert np.all(st == 0)
    print("lu ok", flush=True)


if __name__ == "__main__":
    main()
