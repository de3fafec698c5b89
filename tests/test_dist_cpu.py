"""N > 1 host logic on CPU: world_size 2 over gloo (the GPU path uses NCCL with the same calls).
Instances are partitioned contiguously by rank with no data-path collective; only job totals and
summary statistics are reduced at the end (toy-spice_b200/sharding.py, bench.py)."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_util as PU

S = importlib.import_module("toy-spice_b200.sharding")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the oracle stands in for the per-rank engine here (this test is about the sharding plumbing)
        text = PU.T.BUNDLED["rc"]
        ckt = PU.T.Circuit.from_netlist(text)
        ov_all = PU.draws("rc", ckt, n_total)
        ov = S.shard_overrides(ov_all, rank, world)
        lo, hi = S.shard_range(n_total, rank, world)
        res = PU.O.OracleCircuit(text).run(hi - lo, overrides=ov, want_wave=False, want_stats=True, threads=1)
        stats = res["stats"].transpose(1, 2, 0)                   # -> [4][ncol][n_local], the product's layout
        t_job, counts = S.reduce_job(1.0 + rank, [res["counters"][:, 0].sum(), res["counters"][:, 2].sum()])
        summary = S.merge_summary(stats, res["n_rows"])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), t=t_job, counts=counts, mn=summary["min"], mx=summary["max"],
                 mean=summary["mean"], rows=summary["rows"], lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition():
    for n in (1, 7, 1 << 20, 1000003):
        for w in (1, 2, 4, 8):
            r = [S.shard_range(n, g, w) for g in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[g][1] == r[g + 1][0] for g in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo(tmp_path):
    n_total, world = 24, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 12, 12, 24)
    # every rank ends up with the same job-level numbers
    for k in ("t", "counts", "mn", "mx", "mean", "rows"):
        assert np.array_equal(r0[k], r1[k])
    assert float(r0["t"]) == 2.0                                   # max over ranks
    assert list(r0["counts"]) == [305.0 * n_total, 610.0 * n_total]
    # ... equal to the single-process answer over all instances
    text = PU.T.BUNDLED["rc"]
    ov = PU.draws("rc", PU.T.Circuit.from_netlist(text), n_total)
    res = PU.O.OracleCircuit(text).run(n_total, overrides=ov, want_wave=False, want_stats=True)
    st = res["stats"].transpose(1, 2, 0)
    assert np.array_equal(r0["mn"], st[0].min(axis=1)) and np.array_equal(r0["mx"], st[1].max(axis=1))
    assert np.allclose(r0["mean"], st[2].sum(axis=1) / res["n_rows"].sum(), rtol=1e-13, atol=0)
