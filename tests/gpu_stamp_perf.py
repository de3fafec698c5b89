"""Development driver (not a pytest file): the operator-level two-kernel Newton iteration — device-stamp kernel
(HBM-write-bound: (n^2 + n) * 8 bytes per instance) followed by the warp-per-circuit LU — timed with CUDA events.
Usage: python tests/gpu_stamp_perf.py [deck ...]"""
import json
import os
import sys

import torch

import parity_util as PU

T = PU.T


def main():
    decks = sys.argv[1:] or ["rc", "rlc", "transformer1", "transformer2"]
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    peaks = os.path.join(PU.ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    for name in decks:
        ckt = T.Circuit.from_netlist(T.BUNDLED[name], ctx)
        n = ckt.n
        n_inst = int(min(1 << 24, 4e9 // ((n * n + n) * 8)))
        ov = PU.draws(name, ckt, n_inst)
        b = ckt.batch(n_inst)
        for (d, p), v in ov.items():
            b.set_param(d, p, torch.from_numpy(v).cuda())
        dA = torch.empty((n_inst, n, n), dtype=torch.float64, device="cuda")
        db = torch.empty((n_inst, n), dtype=torch.float64, device="cuda")
        dx = torch.empty_like(db); dst = torch.empty(n_inst, dtype=torch.int32, device="cuda")
        st = ckt.structure()
        order = (st["pivot_row"], st["pivot_col"])
        torch.cuda.synchronize()
        t_stamp, t_lu, t_lu_iso = [], [], []
        flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
        for it in range(4):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            e[0].record(stream)
            b.stamp_dev(T.AN_TRAN, 1e-4, 1e-6, 0.0, dA.data_ptr(), db.data_ptr())
            e[1].record(stream)
            ctx.lu_solve_batched_dev(n, n_inst, dA.data_ptr(), db.data_ptr(), dx.data_ptr(), dst.data_ptr(), order, strict=False)
            e[2].record(stream)
            with torch.cuda.stream(stream):
                flush.zero_()                      # L2 flush, then the LU again on its own
            e[3].record(stream)
            ctx.lu_solve_batched_dev(n, n_inst, dA.data_ptr(), db.data_ptr(), dx.data_ptr(), dst.data_ptr(), order, strict=False)
            e[4].record(stream)
            stream.synchronize()
            if it:
                t_stamp.append(e[0].elapsed_time(e[1])); t_lu.append(e[1].elapsed_time(e[2])); t_lu_iso.append(e[3].elapsed_time(e[4]))
        ts, tl = min(t_stamp) * 1e-3, min(t_lu) * 1e-3
        nvar = len(ov)
        wbytes = n_inst * (n * n + n) * 8
        rbytes = n_inst * nvar * 8
        print(f"{name:13s} n={n:2d} instances={n_inst:9d}  stamp {min(t_stamp):8.3f} ms  {(wbytes + rbytes) / ts / 1e9:7.1f} GB/s "
              f"({(wbytes + rbytes) / ts / 1e9 / hbm * 100:4.1f} % of measured HBM copy peak {hbm:.0f})  {n_inst / ts:.3e} stamps/s   |  "
              f"LU {min(t_lu):8.3f} ms right after the stamp, {min(t_lu_iso):8.3f} ms after an L2 flush  {n_inst / (min(t_lu_iso) * 1e-3):.3e} systems/s  bad={int(dst.sum())}", flush=True)
        del b, dA, db, dx, dst
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
