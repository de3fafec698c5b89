"""Development aid: run bench.py's operator_lu entry alone (python tests/dev/bench_lu_entry.py)."""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
import parity_util as PU  # noqa: E402

T = PU.T
ctx = T.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)


def timed(fn, reps=3, warm=1):
    ms = []
    for i in range(warm + reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream)
        stream.synchronize()
        if i >= warm:
            ms.append(e0.elapsed_time(e1))
    return ms


print(json.dumps(B.measure_lu_operator(T, ctx, timed, "cuda:0", ctx.measure_fp64_peak())))
