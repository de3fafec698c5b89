"""Development aid: the SASS of a kernel from an `ncu --set full --import-source on` report with executed counts and stall
samples per instruction, restricted to instructions executed at least `frac` x the most-executed one (the hot loop).
Usage: python tests/dev/ncu_hot.py <report.ncu-rep> [frac=0.2] [units]   (units: divide counts, e.g. warp-steps)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = rows[1]
ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
body = [r for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
mx = max(int(r[iex]) for r in body)
tot = sum(int(r[iex]) for r in body)
base = int(body[0][ia], 16)
hot = 0
for r in body:
    ex = int(r[iex])
    if ex >= frac * mx:
        hot += ex
        u = f"{ex / units:7.3f}" if units else f"{ex:12d}"
        print(f"{int(r[ia], 16) - base:06x} {u} {int(r[ismp]):6d}  {r[isrc].strip()}")
print(f"# total executed {tot}, listed {hot} ({100.0 * hot / tot:.1f} %)" + (f", per unit {tot / units:.1f}" if units else ""))
