"""Development driver (not a pytest file): every variant of the operator-level batched LU (csrc/lu_warp.cu: rows per lane,
shared-memory pivot-row broadcast [only in a -DTSB_LU_WITH_BCAST=1 build: measured slower everywhere and compiled out, B is
ignored otherwise], asynchronous staging — $TSB_LU_VARIANT = "R,B,A", read per call) in ONE process: the
strict build against the oracle bit for bit, the fast build's backward error, and throughput with CUDA events.
Usage: python tests/gpu_lu_variants.py [bytes_of_A, default 1e9] [n,n,...]"""
import os
import sys

import numpy as np
import torch

import parity_util as PU
from test_lu_operator import mna_like

T, O = PU.T, PU.O
VARIANTS = tuple(os.environ.get("LU_VARIANTS", "1,0,0;1,0,1;2,0,0;2,0,1;3,0,0;3,0,1;4,0,1").split(";"))


def f_lu_dense(n):
    return sum(1 + (n - k) + 2 * (n - k) ** 2 for k in range(1, n + 1)) + n + 2 * n * (n - 1)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
    ns = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (5, 8, 10, 16, 24, 32)
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    print(f"fp64 peak {ctx.measure_fp64_peak():.1f} TFLOP/s", flush=True)
    for n in ns:
        n_inst = int(min(1 << 23, budget // (n * n * 8)))
        base, A1, b1 = mna_like(n, 4096, n)
        order = T.lu_order(base)
        xo, sto, _ = O.lu_batch(base, A1[:515], b1[:515])
        reps = (n_inst + 4095) // 4096
        dA = torch.from_numpy(A1).cuda().repeat(reps, 1, 1)[:n_inst].contiguous()
        db = torch.from_numpy(b1).cuda().repeat(reps, 1)[:n_inst].contiguous()
        dx = torch.empty_like(db)
        dst = torch.empty(n_inst, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        # fixed cost of one call between the two events (permutation upload + launch): a one-system batch
        ms1 = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.lu_solve_batched_dev(n, 1, dA.data_ptr(), db.data_ptr(), dx.data_ptr(), dst.data_ptr(), order, strict=False)
            e1.record(stream)
            stream.synchronize()
            ms1.append(e0.elapsed_time(e1))
        print(f"n={n:2d} one-system call: {min(ms1[1:]):.3f} ms", flush=True)
        for var in VARIANTS:
            os.environ["TSB_LU_VARIANT"] = var
            line = f"n={n:2d} inst={n_inst:8d} [{var}]"
            for strict in (1, 0):
                if strict and var.split(",")[1] == "1":
                    continue                        # the strict build has no shared-memory broadcast: same kernel as "R,0,A"
                dx.fill_(float("nan")); dst.fill_(-1)
                ms = []
                for it in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    ctx.lu_solve_batched_dev(n, n_inst, dA.data_ptr(), db.data_ptr(), dx.data_ptr(), dst.data_ptr(), order, strict=bool(strict))
                    e1.record(stream)
                    stream.synchronize()
                    if it:
                        ms.append(e0.elapsed_time(e1))
                t = min(ms) * 1e-3
                x = dx[:515].cpu().numpy()
                tail = dx[-4096:].cpu().numpy()                  # the last (ragged) pass too
                k0 = (n_inst - 4096) % 4096
                ref_tail = np.roll(np.arange(4096), -k0)         # instance (n_inst - 4096 + i) is a copy of system (k0 + i) % 4096
                bad = int(dst.sum().item())
                if strict:
                    verdict = "bit-identical" if np.array_equal(x, xo) else f"DIFFERS max {np.nanmax(np.abs(x - xo)):.3e} nan={int(np.isnan(x).sum())}"
                else:
                    res = np.einsum("qij,qj->qi", A1[:515], x) - b1[:515]
                    den = np.abs(A1[:515]).sum(axis=2).max(axis=1) * np.abs(x).max(axis=1) + np.abs(b1[:515]).max(axis=1)
                    reso = np.einsum("qij,qj->qi", A1[:515], xo) - b1[:515]
                    verdict = f"berr {np.abs(res).max(axis=1).__truediv__(den).max():.2e} (oracle {(np.abs(reso).max(axis=1) / den).max():.2e})"
                # the tail against the head: same systems, same bits (every pass of the grid-stride loop computes alike)
                head_all = dx[:4096].cpu().numpy()
                same_tail = np.array_equal(tail, head_all[ref_tail])
                line += f" | strict={strict} {min(ms):7.3f} ms {n_inst / t:.3e}/s {n_inst * f_lu_dense(n) / t / 1e12:5.2f} TF/s {n_inst * ((n * n + 2 * n) * 8 + 4) / t / 1e9:6.0f} GB/s {verdict} tail={'ok' if same_tail else 'BAD'} bad={bad}"
            print(line, flush=True)
        os.environ.pop("TSB_LU_VARIANT", None)
        del dA, db, dx, dst
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
