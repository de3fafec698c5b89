"""Multi-GPU plumbing: one process per GPU, instances partitioned contiguously by rank, NO collective
on the data path.  The only exchanges are end-of-run reductions of a few numbers
(job totals, per-signal summary statistics) through torch.distributed (NCCL on GPUs, gloo in the CPU
tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous instance range [g*N/G, (g+1)*N/G) of rank g."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def shard_overrides(overrides: dict, rank: int, world: int) -> dict:
    """Slice every per-instance parameter array to this rank's instance range."""
    out = {}
    for k, v in overrides.items():
        lo, hi = shard_range(len(v), rank, world)
        out[k] = np.ascontiguousarray(v[lo:hi])
    return out


def reduce_job(t_local: float, counts_local, device=None):
    """(max over ranks of the elapsed time, sum over ranks of the counters).  Timing of a multi-GPU job is
    the slowest rank's device time; work is the sum of all ranks' accepted steps / solves."""
    import torch
    import torch.distributed as dist
    counts = [float(c) for c in counts_local]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(t_local), counts
    t = torch.tensor([t_local], dtype=torch.float64, device=device)
    c = torch.tensor(counts, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t[0]), [float(x) for x in c]


def merge_summary(stats_local: np.ndarray, rows_local: np.ndarray, device=None) -> dict:
    """Global per-column summary over ALL instances of ALL ranks from the per-instance statistics
    [4: min, max, sum, last][ncol][n_local]: global min, global max, mean over every stored row."""
    import torch
    import torch.distributed as dist
    mn = stats_local[0].min(axis=1)
    mx = stats_local[1].max(axis=1)
    sm = stats_local[2].sum(axis=1)
    cnt = float(rows_local.sum())
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        tmn = torch.tensor(mn, dtype=torch.float64, device=device)
        tmx = torch.tensor(mx, dtype=torch.float64, device=device)
        tsm = torch.tensor(np.concatenate([sm, [cnt]]), dtype=torch.float64, device=device)
        dist.all_reduce(tmn, op=dist.ReduceOp.MIN)
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
        mn, mx = tmn.cpu().numpy(), tmx.cpu().numpy()
        sm, cnt = tsm[:-1].cpu().numpy(), float(tsm[-1])
    return dict(min=mn, max=mx, mean=sm / max(cnt, 1.0), rows=cnt)
