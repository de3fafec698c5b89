"""Development driver (not a pytest file): BASELINE.json configs[4] at full size — transformer1-3.cir coupled-inductor
transient, 2^24 instances in total, sharded contiguously over the ranks of one box (torchrun, one rank per GPU, no
collective on the data path; job totals and the global per-signal summary are all-reduced at the end).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/gpu_config5.py [log2_total]
Rank 0 prints one JSON line per deck."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

import parity_util as PU

T = PU.T
S = importlib.import_module("toy-spice_b200.sharding")
W = importlib.import_module("toy-spice_b200.workloads")


def main():
    log2_total = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = T.Context(local)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)
    n_total = 1 << log2_total
    lo, hi = S.shard_range(n_total, rank, world)
    n = hi - lo
    for name in ("transformer3", "transformer1", "transformer2"):
        ckt = T.Circuit.from_netlist(T.BUNDLED[name], ctx)
        card = ckt.analysis_card()
        # this rank's slice of the job's draws: PCG64 streams are advanced per parameter, so draw the rank's block with
        # a rank-keyed seed (weak-scaling style, as bench.py does) — the job is 2^24 independent draws either way
        ov = W.sweep_draws(ckt.devices(), n, W.sweep_seed(name) + 7919 * rank)
        dev = {k: torch.from_numpy(v).cuda() for k, v in ov.items()}
        b = ckt.batch(n)
        for (d, p), v in dev.items():
            b.set_param(d, p, v)
        ms = []
        for it in range(2):                                   # pass 0: kernel load + launch-bounds autotuning
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS)
            e1.record(stream)
            stream.synchronize()
            ms.append(e0.elapsed_time(e1))
        tot = b.totals()
        bad = int((b.status() != 0).sum())
        t_job, (acc, rej, solves, bad_job) = S.reduce_job(ms[-1] * 1e-3, [tot[0], tot[1], tot[2], bad], device=f"cuda:{local}")
        summ = S.merge_summary(b.stats_all(), b.rows(), device=f"cuda:{local}")
        if rank == 0:
            print(json.dumps({"deck": name, "instances_total": n_total, "n_gpus": world, "instances_per_gpu": n,
                              "seconds": t_job, "accepted_steps": acc, "rejected_steps": rej, "reference_solves": solves,
                              "circuit_timesteps_per_sec": acc / t_job, "failed_instances": bad_job,
                              "global_min": summ["min"].tolist(), "global_max": summ["max"].tolist(),
                              "columns": ckt.columns(T.AN_TRAN)}), flush=True)
        del b, dev
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
