"""Turns an `ncu --set full --import-source on` report into the small CSV summaries committed in this directory.
Usage: python profiles/summarize_ncu.py <report.ncu-rep> <out.csv> "<label>" [units_per_launch]
(units = accepted steps or systems per launch in warps, for the per-unit instruction mix)."""
import collections
import csv
import re
import subprocess
import sys

METRICS = """gpu__time_duration.sum launch__registers_per_thread launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
launch__waves_per_multiprocessor sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sass__inst_executed_local_loads
sass__inst_executed_local_stores sm__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed smsp__warps_eligible.avg.per_cycle_active""".split()


def page(rep, which):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout.splitlines()))


def main():
    rep, out, label = sys.argv[1:4]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = page(rep, "raw")
    hdr, unit_row, vals = raw[0], raw[1], raw[2]
    d = {h: (u, v) for h, u, v in zip(hdr, unit_row, vals)}
    rows = [["metric", "unit", label]]
    for k in ("Kernel Name", "Grid Size", "Block Size"):
        if k in d:
            rows.append([k, "", d[k][1]])
    for k in METRICS + sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")):
        if k in d:
            rows.append([k, d[k][0], d[k][1]])
    src = page(rep, "source")
    h2 = src[1]
    isrc, iex = h2.index("Source"), h2.index("Instructions Executed")
    op = collections.Counter()
    for r in src[2:]:
        try:
            ex = int(r[iex])
        except (ValueError, IndexError):
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
        op[t.split()[0].split(".")[0]] += ex
    tot = sum(op.values())
    rows.append(["executed warp instructions by opcode (top 14)", "% of all" + (", per unit" if units else ""), ""])
    for k, v in op.most_common(14):
        rows.append(["  " + k, f"{100.0 * v / tot:.1f}", f"{v / units:.2f}" if units else ""])
    if units:
        rows.append(["warp instructions per unit", "", f"{tot / units:.1f}"])
    csv.writer(open(out, "w")).writerows(rows)
    print(out, len(rows), "rows")


if __name__ == "__main__":
    main()
