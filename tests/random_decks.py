"""Seeded generator of small random SPICE decks (R, C, L, D, V, I; DC / SIN / PULSE / PWL sources; mixed-case
names, engineering suffixes) used to exercise the front-end, the symbolic pass, the code generator and the
analysis kernels beyond the bundled decks.  Every node has a resistive path to ground."""
import numpy as np


def _eng(v: float, rng) -> str:
    """A value the reference's ParseValue accepts, in a randomly chosen spelling."""
    for suf, mul in (("meg", 1e6), ("k", 1e3), ("", 1.0), ("m", 1e-3), ("u", 1e-6), ("n", 1e-9), ("p", 1e-12)):
        if v >= mul:
            mant = v / mul
            style = rng.integers(0, 3)
            if style == 0:
                return f"{mant:.4g}{suf}"
            if style == 1:
                return f"{mant:.3f}{'K' if suf == 'k' else suf}"      # the reference's suffix set is case-sensitive: K == k only
            return f"{v:.6e}"
    return f"{v:.6e}"


def random_deck(seed: int, allow_inductor=True, allow_diode=True):
    rng = np.random.default_rng(seed)
    m = int(rng.integers(2, 6))                     # nodes 1..m
    lines = [f"* random deck {seed}"]
    src = rng.choice(["sin", "pulse", "dc", "pwl"])
    tstop, tstep = 2e-3, 2e-5
    if src == "sin":
        lines.append(f"Vsrc 1 0 SIN({rng.uniform(-1, 1):.3f} {rng.uniform(1, 8):.3f} {rng.choice([500, 1000, 2500])})")
    elif src == "pulse":
        lines.append("Vsrc 1 0 PULSE(0 5 0.1m 0.05m 0.08m 0.4m 1m)")
    elif src == "pwl":
        lines.append("Vsrc 1 0 PWL(0 0 0.2m 0 0.5m 3.3 1m 3.3 1.2m -1 2m 0)")
    else:
        lines.append(f"Vsrc 1 0 DC {rng.uniform(1, 12):.3f}")
    cnt = {"R": 0, "C": 0, "L": 0, "D": 0, "I": 0}

    def name(k):
        cnt[k] += 1
        base = f"{k}{cnt[k]}"
        return base.lower() if rng.random() < 0.3 else base

    for k in range(2, m + 1):                       # spanning chain: node k hangs on an earlier node through a resistor
        other = int(rng.integers(1, k))
        lines.append(f"{name('R')} {other} {k} {_eng(float(np.exp(rng.uniform(np.log(50), np.log(2e4)))), rng)}")
    lines.append(f"{name('R')} {m} 0 {_eng(float(np.exp(rng.uniform(np.log(100), np.log(1e4)))), rng)}")
    has_l = False
    for _ in range(int(rng.integers(1, 4))):
        kind = rng.choice(["C", "R", "L", "D", "I"], p=[0.35, 0.2, 0.15, 0.2, 0.1])
        a = int(rng.integers(1, m + 1))
        b = int(rng.integers(0, m + 1))
        if b == a:
            b = 0
        if kind == "C":
            lines.append(f"{name('C')} {a} {b} {_eng(float(np.exp(rng.uniform(np.log(1e-8), np.log(2e-6)))), rng)}")
        elif kind == "R":
            lines.append(f"{name('R')} {a} {b} {_eng(float(np.exp(rng.uniform(np.log(100), np.log(1e5)))), rng)}")
        elif kind == "L" and allow_inductor and not has_l and a != 1 and b != 1:
            has_l = True
            lines.append(f"{name('L')} {a} {b} {_eng(float(np.exp(rng.uniform(np.log(1e-4), np.log(1e-2)))), rng)}")
        elif kind == "D" and allow_diode:
            lines.append(f"{name('D')} {a} {b} D")
        elif kind == "I":
            lines.append(f"{name('I')} {a} 0 DC {rng.uniform(0.1, 2):.3f}m")
    if has_l:
        tstop, tstep = 4e-4, 4e-6                   # inductor decks take ~50 x 300 steps whatever the span
    lines.append(f".tran {tstep:g} {tstop:g}")
    return "\n".join(lines) + "\n", dict(nodes=m, has_inductor=has_l, source=str(src))


def rc_ladder(sections: int) -> str:
    """Vin - (R - C to ground) x sections: n = sections + 2 unknowns, tridiagonal conductance block (no fill)."""
    lines = [f"* RC ladder, {sections} sections", "Vin 1 0 SIN(0 5 1k)"]
    for k in range(1, sections + 1):
        lines.append(f"R{k} {k} {k + 1} 100")
        lines.append(f"C{k} {k + 1} 0 100n")
    lines.append(".tran 0.01ms 3ms")
    return "\n".join(lines) + "\n"
