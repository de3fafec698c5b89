"""Development driver (not a pytest file): throughput of the operator-level batched LU (csrc/lu_warp.cu), device
pointers, CUDA events.  Usage: python tests/gpu_lu_perf.py [bytes_of_A, default 2e9] [n,n,...]"""
import sys

import numpy as np
import torch

import parity_util as PU
from test_lu_operator import mna_like

T = PU.T


def f_lu_dense(n):
    return sum(1 + (n - k) + 2 * (n - k) ** 2 for k in range(1, n + 1)) + n + 2 * n * (n - 1)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 2e9
    ctx = T.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    peak = ctx.measure_fp64_peak()
    print(f"fp64 peak {peak:.1f} TFLOP/s")
    ns = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (3, 5, 8, 10, 16, 24, 32)
    for n in ns:
        n_inst = int(min(1 << 24, budget // (n * n * 8)))
        base, A1, b1 = mna_like(n, 4096, n)
        order = T.lu_order(base)
        reps = (n_inst + 4095) // 4096
        dA = torch.from_numpy(A1).cuda().repeat(reps, 1, 1)[:n_inst].contiguous()
        db = torch.from_numpy(b1).cuda().repeat(reps, 1)[:n_inst].contiguous()
        dx = torch.empty_like(db)
        dst = torch.empty(n_inst, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for strict in (0, 1):
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
            byts = n_inst * ((n * n + 2 * n) * 8 + 4)
            print(f"n={n:2d} inst={n_inst:9d} strict={strict}  {min(ms):8.3f} ms  {n_inst / t:.3e} solves/s  {byts / t / 1e9:7.1f} GB/s  "
                  f"{n_inst * f_lu_dense(n) / t / 1e12:6.2f} TFLOP/s (dense count {f_lu_dense(n)})  bad={int(dst.sum())}", flush=True)
        del dA, db, dx, dst
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
