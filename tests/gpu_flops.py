"""Development driver (not a pytest file): the target of an `ncu --metrics <FP64 opcode counts>` pass that measures the FP64
arithmetic every bench deck ACTUALLY executes per solve (the generated code skips structural zeros, condenses invariant
pivots and hoists invariant stamps, so the dense-algorithm flop model of SURVEY §8(d) overstates it).

    ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/flops_ncu.csv python tests/gpu_flops.py > gpurun_out/flops_runs.jsonl
    python profiles/flops_from_ncu.py gpurun_out/flops_ncu.csv gpurun_out/flops_runs.jsonl profiles/executed_flops.json

Every deck runs with its own instance count (65536 + 256*k) so that its launch is recognisable in the launch list by
grid size x block size; the LAST matching launch is the measured one (after the priming run)."""
import json
import sys

import torch

import parity_util as PU

T = PU.T
TRAN150 = dict(tstart=0.0, tstop=150e-6, tstep=1e-6, tmax=0.0, uic=False)
DECKS = [("rc", None), ("rlc", None), ("rl", None), ("diode2", None), ("diode4", None), ("diode1", None), ("diode5", None), ("diode3", None),
         ("mosfet1", None), ("bjt2", None), ("bjt1", TRAN150), ("bjt3", TRAN150), ("transformer3", None), ("transformer1", None),
         ("transformer2", None)]


def main():
    strict = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ctx = T.Context(0)
    for k, (deck, tran) in enumerate(DECKS):
        n = 65536 + 256 * k
        ckt = T.Circuit.from_netlist(T.BUNDLED[deck], ctx)
        card = ckt.analysis_card()
        if tran:
            card = dict(card, **tran, analysis=T.AN_TRAN)
        an = card["analysis"]
        ov = PU.draws(deck, ckt, n)
        keep = {key: torch.from_numpy(v).cuda() for key, v in ov.items()}
        b = ckt.batch(n)
        for (d, p), v in keep.items():
            b.set_param(d, p, v)
        o = T.default_opts(strict_fp=strict, share_time_grid=1)      # the bench's batches (>= 2^18 instances) run the shared-time-grid kernels
        for _ in range(2):
            if an == T.AN_TRAN:
                b.run_tran(card["tstart"], card["tstop"], card["tstep"], card["tmax"], card["uic"], out=T.OUT_STATS, opts=o)
            elif an == T.AN_OP:
                b.run_op(o)
            else:
                b.run_dc(card["dc_src_dev"], card["dc_start"], card["dc_stop"], card["dc_inc"], out=T.OUT_STATS, opts=o)
            b.sync()
        tot = b.totals()
        print(json.dumps({"deck": deck, "analysis": {T.AN_OP: "op", T.AN_TRAN: "tran", T.AN_DC: "dc"}[an], "instances": n, "n": ckt.n,
                          "strict_fp": strict, "kernel_key": b.kernel_key(o), "accepted_steps": int(tot[0]), "rejected_steps": int(tot[1]),
                          "reference_solves": int(tot[2] + tot[3]), "executed_solves": int(tot[4])}), flush=True)
        del b, keep


if __name__ == "__main__":
    main()
