#!/usr/bin/env python
"""bench.py — circuit-timesteps/sec of the batched FP64 transient hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): rc.cir + rlc.cir linear transient, 2^20-instance R/L/C parameter
sweep per GPU (nominal * LogUniform[0.5, 2], PCG64 seed 1234), summary-statistics
output (min / max / sum / last per signal + row counts; full rlc waveforms for 2^20 instances would be
~640 GB).  One "step" = one pass of the hot path over that batch: one transient launch per deck.

  value     whole-job accepted transient steps / s with parameters already resident in HBM
  e2e       the same through the public API with HOST buffers: per step the parameter arrays are
            copied H2D from pinned memory and the statistics / row counts / status are read back
  roofline  the dominant launch (rlc transient): algorithmic FP64 flops / CUDA-event time vs the
            DFMA-chain peak measured in this run (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline / --impl reference
            the CPU restatement of the reference solver (oracle/, `kind: port` — the Go reference
            cannot be built here), one task per instance on all host threads, bounded sample.

Launch: `python bench.py --gpus N --steps K --warmup W`; for N > 1 under torchrun (one rank per GPU,
instances sharded by rank, no collective on the data path; totals all-reduced at the end).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DECKS = ("rc", "rlc")


# --------------------------------------------------------------------------------------------
# Algorithmic FP64 work per Newton solve, stated in DESIGN.md §Roofline.
def f_lu_dense(n: int) -> int:
    """Dense LU + triangular solves as the reference performs them (structure is dense n x n,
    SURVEY Q16): sum_k [1 + (n-k) + 2(n-k)^2] + n + 2n(n-1)."""
    return sum(1 + (n - k) + 2 * (n - k) ** 2 for k in range(1, n + 1)) + n + 2 * n * (n - 1)


STAMP_FLOPS = {0: 5, 1: 8, 2: 9, 3: 5, 4: 2, 5: 30, 6: 90, 7: 80, 8: 12, 9: 9}   # per device kind (adds into A/b + model arithmetic)


def ncu_traffic(deck: str, n: int, mode: str):
    """DRAM bytes per launch of this kernel from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{deck}:{n}:{mode}")
    except Exception:
        return None


def ncu_counters(deck: str, n: int, mode: str):
    """FP64-pipe / issue utilisation of this kernel from the committed ncu capture, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"ncu:{deck}:{n}:{mode}")
    except Exception:
        return None


def flops_per_solve(ckt) -> dict:
    n = ckt.n
    stamp = sum(STAMP_FLOPS[d["kind"]] for d in ckt.devices())
    conv = 5 * n
    return dict(n=n, lu=f_lu_dense(n), stamp=stamp, conv=conv, total=f_lu_dense(n) + stamp + conv)


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.rows:
            if ts < t0 - 0.1 or ts > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); smax = max(smax, float(f[2]))
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------
def cpu_sample(T, O, W, target_seconds: float, threads: int):
    """Time the CPU oracle on a bounded sample of the SAME workload (same decks, same draws)."""
    per_deck = {}
    total_steps, total_time = 0, 0.0
    desc = []
    for name in DECKS:
        oc = O.OracleCircuit(T.BUNDLED[name])
        ckt = T.Circuit.from_netlist(T.BUNDLED[name])
        probe_n = max(threads, 8)
        ov = W.sweep_draws(ckt.devices(), probe_n, W.sweep_seed(name))
        t0 = time.time()
        oc.run(probe_n, overrides=ov, threads=threads, want_wave=False, want_stats=True)
        rate = probe_n / max(time.time() - t0, 1e-6)
        n = int(min(1 << 16, max(probe_n, rate * target_seconds / len(DECKS))))
        ov = W.sweep_draws(ckt.devices(), n, W.sweep_seed(name))
        t0 = time.time()
        res = oc.run(n, overrides=ov, threads=threads, want_wave=False, want_stats=True)
        dt = time.time() - t0
        steps = int(res["counters"][:, 0].sum())
        per_deck[name] = dict(instances=n, seconds=dt, steps=steps)
        # weight decks as in the full workload (equal instance counts): time per instance
        total_steps += steps / n
        total_time += dt / n
        desc.append(f"{name}: {n} instances in {dt:.2f}s")
    return total_steps / total_time, "; ".join(desc), per_deck


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Go solver cannot be
    built in this image (no Go toolchain, un-vendored sparse module), so this is the C++ restatement
    (oracle/, kind 'port') on all host threads, one task per instance, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    T = importlib.import_module("toy-spice_b200")
    W = importlib.import_module("toy-spice_b200.workloads")
    from oracle import oracle as O
    threads = host_threads()
    for _ in range(args.warmup):
        cpu_sample(T, O, W, 1.0, threads)
    vals, t_all, sample = [], 0.0, ""
    for _ in range(args.steps):
        t0 = time.time()
        v, sample, _ = cpu_sample(T, O, W, max(2.0, 20.0 / max(1, args.steps)), threads)
        t_all += time.time() - t0
        vals.append(v)
    value = float(np.mean(vals))
    line = {
        "metric": "circuit-timesteps/sec (batched FP64 transient)", "value": value, "unit": "circuit-timesteps/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_all / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "rc.cir + rlc.cir transient, R/L/C sweep (LogUniform[0.5,2], PCG64 seed 1234), bounded CPU sample",
                   "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "circuit-timesteps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "circuit-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=1 << 20, help="instances per deck per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strict-fp", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T = importlib.import_module("toy-spice_b200")
    W = importlib.import_module("toy-spice_b200.workloads")
    ctx = T.Context(local)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)
    opts = T.default_opts(strict_fp=args.strict_fp)
    n = args.instances

    # ---- workload: decks, per-rank parameter draws (rank r owns its own instance range) ---------
    decks = []
    h2d_bytes = 0
    for name in DECKS:
        ckt = T.Circuit.from_netlist(T.BUNDLED[name], ctx)
        card = ckt.analysis_card()
        ov = W.sweep_draws(ckt.devices(), n, W.sweep_seed(name) + 7919 * rank)
        host = {k: torch.from_numpy(v).pin_memory() for k, v in ov.items()}
        dev = {k: v.to(f"cuda:{local}", non_blocking=True) for k, v in host.items()}
        b_res = ckt.batch(n)       # parameters resident in HBM (borrowed torch tensors)
        for (d, p), v in dev.items():
            b_res.set_param(d, p, v)
        b_e2e = ckt.batch(n)       # parameters copied from pinned host memory every step
        for (d, p), v in host.items():
            b_e2e.set_param(d, p, v.numpy())
        h2d_bytes += sum(v.numel() * 8 for v in host.values())
        decks.append(dict(name=name, ckt=ckt, card=card, host=host, dev=dev, b_res=b_res, b_e2e=b_e2e,
                          flops=flops_per_solve(ckt), ncol=len(ckt.columns(T.AN_TRAN))))
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=f"cuda:{local}")   # 256 MB > 126 MB L2

    def launch_deck(d, batch):
        c = d["card"]
        batch.run_tran(c["tstart"], c["tstop"], c["tstep"], c["tmax"], c["uic"], out=T.OUT_STATS, opts=opts)

    def step_resident():
        for d in decks:
            launch_deck(d, d["b_res"])

    d2h_bytes = sum(n * (4 * d["ncol"] * 8 + 8 + 4) for d in decks)
    host_out = {d["name"]: dict(stats=torch.empty((4, d["ncol"], n), dtype=torch.float64).pin_memory().numpy(),
                                rows=torch.empty(n, dtype=torch.int64).pin_memory().numpy(),
                                status=torch.empty(n, dtype=torch.int32).pin_memory().numpy()) for d in decks}

    def step_e2e():
        for d in decks:
            b = d["b_e2e"]
            for (dv, p), v in d["host"].items():
                b.set_param(dv, p, v.numpy())          # cudaMemcpyAsync H2D from pinned memory
            launch_deck(d, b)
        for d in decks:
            b = d["b_e2e"]
            o = host_out[d["name"]]
            b.stats_all(o["stats"])
            b.rows(o["rows"])
            b.status(o["status"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- FP64 roof of this GPU -----------------------------------------------------------------
    fp64_peak = ctx.measure_fp64_peak()

    # ---- warm-up (also loads / compiles the specialised kernels) --------------------------------
    step_resident()             # priming pass (not one of the W warm-up steps): loads the kernels, runs the launch-bounds autotuner
    for _ in range(max(args.warmup, 3 if args.warmup > 0 else 0)):
        step_resident()
    torch.cuda.synchronize()

    # ---- timed: resident parameters (value), per-step CUDA events on the launch stream ----------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    launches0 = ctx.launch_count
    t_wall0 = time.time()
    ms_steps, ms_dom = [], []
    for _ in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(len(decks) + 1)]
        with torch.cuda.stream(stream):
            e[0].record(stream)
            for i, d in enumerate(decks):
                launch_deck(d, d["b_res"])
                e[i + 1].record(stream)
        stream.synchronize()
        ms_steps.append(e[0].elapsed_time(e[-1]))
        ms_dom.append(e[len(decks) - 1].elapsed_time(e[len(decks)]))      # last deck = rlc = dominant launch
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(t_wall0, t_wall1)

    totals = {d["name"]: d["b_res"].totals() for d in decks}
    acc_local = int(sum(t[0] for t in totals.values()))
    solves_local = int(sum(t[2] + t[3] for t in totals.values()))     # as the reference counts them
    exec_local = int(sum(t[4] for t in totals.values()))              # factor+solve passes actually executed
    t_local = sum(ms_steps) * 1e-3
    S = importlib.import_module("toy-spice_b200.sharding")
    t_job, (acc_job, solves_job, exec_job) = S.reduce_job(t_local, [acc_local, solves_local, exec_local], device=f"cuda:{local}")
    value = acc_job * args.steps / t_job if t_job > 0 else 0.0

    # ---- timed: end to end through the public API with host buffers -----------------------------
    step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e_local = time.time() - t0
    te = torch.tensor([t_e2e_local], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = acc_job * args.steps / float(te[0])
    bad_status = int(sum((host_out[d["name"]]["status"] != 0).sum() for d in decks))

    # ---- roofline of the dominant launch (rlc transient) -----------------------------------------
    dom = decks[-1]
    dom_tot = totals[dom["name"]]
    dom_solves = int(dom_tot[4])          # EXECUTED factor+solve passes (the redundant linear re-solve is not run, not counted)
    dom_flops = dom_solves * dom["flops"]["total"]
    dom_ms = float(np.mean(ms_dom))
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp64", "kernel": "tsb_optran (rlc.cir)", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved / fp64_peak if fp64_peak else None, "traffic": ncu_traffic(dom["name"], n, "stats"),
        "peak_source": "DFMA-chain microbenchmark measured in this run (tsb_ctx_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
        "flops_per_solve": dom["flops"], "executed_solves_per_launch": dom_solves, "ms_per_launch": dom_ms,
        # `achieved` / `frac` count only the factor+solve passes this kernel EXECUTES (the conservative reading).
        # SURVEY §8(d)'s per-unit figure is iters*(F_LU + F_stamp + F_conv) per circuit-timestep with iters = 2 for a
        # linear circuit, i.e. the work of the reference algorithm, whose second solve and LTE-rejected solves this
        # kernel proves redundant and does not run: reported beside it, not instead of it.
        "reference_algorithm": {
            "flops_per_launch": int(dom_tot[2]) * dom["flops"]["total"], "solves_per_launch": int(dom_tot[2]),
            "achieved": int(dom_tot[2]) * dom["flops"]["total"] / (dom_ms * 1e-3) / 1e12,
            "frac": (int(dom_tot[2]) * dom["flops"]["total"] / (dom_ms * 1e-3) / 1e12) / fp64_peak if fp64_peak else None},
        "ncu": ncu_counters(dom["name"], n, "stats"),
        "algorithmic_hbm_bytes_per_launch": n * (8 * 3 + 4 * dom["ncol"] * 8 + 8 + 4 + 6 * 8),
    }

    # ---- the HBM-bound regime of the same kernel: rc.cir with every stored row materialised --------------
    # (8*O bytes per stored row; not part of `value`, reported beside the FP64 roofline)
    roofline_hbm = None
    try:
        rc = decks[0]
        c = rc["card"]
        cap = 320
        bw = rc["ckt"].batch(n)
        for (d, p), v in rc["dev"].items():
            bw.set_param(d, p, v)
        ms_w = []
        for i in range(4):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            bw.run_tran(c["tstart"], c["tstop"], c["tstep"], c["tmax"], c["uic"], out=T.OUT_WAVE, cap_rows=cap, opts=opts)
            e1.record(stream)
            stream.synchronize()
            if i > 0:
                ms_w.append(e0.elapsed_time(e1))
        rows_total = int(bw.totals()[0])                    # rc stores every accepted step (305 rows per instance)
        alg_bytes = rows_total * rc["ncol"] * 8 + n * (2 * 8 + 8 + 4 + 8 * 8)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            hbm_peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        ach = alg_bytes / (min(ms_w) * 1e-3) / 1e9
        roofline_hbm = {"bound": "hbm", "kernel": "tsb_optran (rc.cir, TSB_OUT_WAVE)", "achieved": ach, "peak": hbm_peak,
                        "unit": "GB/s", "frac": ach / hbm_peak, "traffic": ncu_traffic("rc", n, "wave"), "peak_source": src,
                        "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": min(ms_w),
                        "circuit_timesteps_per_sec": rows_total / (min(ms_w) * 1e-3)}
        del bw
    except Exception as ex:       # the FP64-bound headline must not depend on this extra measurement
        roofline_hbm = {"error": str(ex)[:200]}

    # ---- the operator-level device-stamp kernel (HBM-write-bound: (n^2 + n) * 8 bytes per instance) -----------------
    # rlc.cir transient companion stamps of 2^22 instances, coalesced through shared memory (tsb_stamp_staged); not part
    # of `value` (the analysis kernels never materialise the matrix), reported as the stamp stage's own roofline
    roofline_stamp = None
    try:
        ns = 1 << 22
        rl = decks[-1]
        nn = rl["ckt"].n
        ovs = W.sweep_draws(rl["ckt"].devices(), ns, 99)
        bs = rl["ckt"].batch(ns)
        keep = []
        for (d, p), v in ovs.items():
            t = torch.from_numpy(v).to(f"cuda:{local}"); keep.append(t)
            bs.set_param(d, p, t)
        dA = torch.empty((ns, nn, nn), dtype=torch.float64, device=f"cuda:{local}")
        dB = torch.empty((ns, nn), dtype=torch.float64, device=f"cuda:{local}")
        ms_s = []
        for i in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            bs.stamp_dev(T.AN_TRAN, 1e-4, 1e-6, 0.0, dA.data_ptr(), dB.data_ptr(), opts=opts)
            e1.record(stream)
            stream.synchronize()
            if i > 0:
                ms_s.append(e0.elapsed_time(e1))
        sbytes = ns * ((nn * nn + nn) * 8 + len(ovs) * 8)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hp, hsrc = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            hp, hsrc = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        ach_s = sbytes / (min(ms_s) * 1e-3) / 1e9
        roofline_stamp = {"bound": "hbm", "kernel": "tsb_stamp_staged (rlc.cir transient stamps, operator level)", "achieved": ach_s, "peak": hp,
                          "unit": "GB/s", "frac": ach_s / hp, "traffic": None, "peak_source": hsrc, "algorithmic_bytes_per_launch": sbytes,
                          "ms_per_launch": min(ms_s), "instances": ns}
        del bs, dA, dB, keep
    except Exception as ex:
        roofline_stamp = {"error": str(ex)[:200]}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        threads = host_threads()
        v, sample, _ = cpu_sample(T, O, W, 16.0, threads)
        cpu = {"value": v, "unit": "circuit-timesteps/s", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "circuit-timesteps/sec (batched FP64 transient)", "value": value, "unit": "circuit-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_job / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"rc.cir + rlc.cir transient, {n}-instance R/L/C sweep per GPU (LogUniform[0.5,2], PCG64 seed 1234), "
                                   "summary-statistics output", "instances_per_deck_per_gpu": n, "parallelism": f"instances sharded over {world} GPU(s), no collective",
                       "l2": "flushed between timed iterations (256 MB write)", "strict_fp": args.strict_fp},
            "accepted_steps_per_step": acc_job, "newton_solves_per_step": solves_job,
            "newton_solves_per_sec": solves_job * args.steps / t_job if t_job > 0 else 0.0,
            "executed_solves_per_step": exec_job,
            "e2e": {"value": e2e_value, "unit": "circuit-timesteps/s", "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes * world},
            "gpu_launches": int(launches), "failed_instances": bad_status,
            "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_hbm_stamp": roofline_stamp, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
