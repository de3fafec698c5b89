"""Joins the launch list of `ncu --metrics <FP64 opcode counts>` over tests/gpu_flops.py with that script's own output:
executed FP64 flops (2*DFMA + DMUL + DADD, thread level, predicated-on) per executed solve and per accepted step of every
bench deck -> profiles/executed_flops.json, which bench.py uses for `roofline.frac` (executed arithmetic over the measured
FP64 peak) instead of the dense-algorithm model.
Usage: python profiles/flops_from_ncu.py <ncu.csv> <runs.jsonl> <out.json>"""
import csv
import json
import sys


def main():
    ncu_csv, runs, out = sys.argv[1:4]
    lines = [l for l in open(ncu_csv).read().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    launches = {}                       # launch ID -> dict
    for r in rows:
        L = launches.setdefault(r["ID"], {"kernel": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        L[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        L[r["Metric Name"] + ":unit"] = r["Metric Unit"]
    def threads(L):
        g = [int(x) for x in L["grid"].strip("()").split(",")]
        b = [int(x) for x in L["block"].strip("()").split(",")]
        return g[0] * g[1] * g[2] * b[0] * b[1] * b[2], b[0]
    table = {}
    for line in open(runs):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        n = d["instances"]
        match = None
        for lid in sorted(launches, key=int):
            L = launches[lid]
            t, blk = threads(L)
            if L["kernel"].startswith("tsb_") and n <= t < n + blk:
                match = L
        if match is None:
            continue
        dfma = match["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
        dmul = match["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
        dadd = match["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
        flops = 2 * dfma + dmul + dadd
        dur = match["gpu__time_duration.sum"]
        unit = match["gpu__time_duration.sum:unit"]
        ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        key = d["deck"] + (":strict" if d["strict_fp"] else "")
        table[key] = dict(d, kernel=match["kernel"], grid=match["grid"], block=match["block"], dfma=dfma, dmul=dmul, dadd=dadd, flops=flops,
                          flops_per_executed_solve=flops / max(1, d["executed_solves"]),
                          flops_per_accepted_step=(flops / d["accepted_steps"]) if d["analysis"] == "tran" and d["accepted_steps"] else None,
                          warp_instructions=match.get("smsp__inst_executed.sum"),
                          warp_instructions_per_executed_solve=32 * match.get("smsp__inst_executed.sum", 0) / max(1, d["executed_solves"]),
                          ncu_ms=ms)
    json.dump({"what": "executed FP64 flops per deck from ncu opcode counts (tests/gpu_flops.py); 2*DFMA + DMUL + DADD, thread-level, predicated-on",
               "decks": table}, open(out, "w"), indent=1)
    for k, v in table.items():
        print(f"{k:22s} n={v['n']:2d} flops/solve={v['flops_per_executed_solve']:8.1f} warp-instr/solve(lane)={v['warp_instructions_per_executed_solve']:8.1f} ms={v['ncu_ms']:.3f}")


if __name__ == "__main__":
    main()
