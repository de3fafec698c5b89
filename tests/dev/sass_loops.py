"""Development aid (not a test): list the loops of one kernel of a cubin from `cuobjdump -sass` — backward branches with the
opcode histogram of the instructions between target and branch — so that instruction counts per loop trip can be checked
here, without GPU time.  Usage: python tests/dev/sass_loops.py <cubin> [kernel] [--dump LO HI]"""
import collections
import re
import subprocess
import sys


def parse(cubin, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    ins, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = line.strip().endswith(": " + kernel)
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def main():
    cubin = sys.argv[1]
    kernel = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "tsb_optran"
    ins = parse(cubin, kernel)
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    if "--dump" in sys.argv:
        k = sys.argv.index("--dump")
        lo, hi = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
        for a, t in ins:
            if lo <= a <= hi:
                print(f"{a:06x}  {t}")
        return
    print(f"{kernel}: {len(ins)} instructions")
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)\)?", t)
        if m and opcode(t) == "BRA":
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx:
                body = ins[addr_idx[tgt]:i + 1]
                h = collections.Counter(opcode(x) for _, x in body)
                top = ", ".join(f"{k} {v}" for k, v in h.most_common(14))
                print(f"loop {tgt:06x}..{a:06x}: {len(body)} instr  [{t}]\n     {top}")


if __name__ == "__main__":
    main()
