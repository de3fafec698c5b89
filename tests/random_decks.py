"""Seeded generators of small random SPICE decks: random_deck (R, C, L, D, V, I; DC / SIN / PULSE / PWL sources; mixed-case
names, engineering suffixes) and random_active_deck (Q, M, K, core inductors) used to exercise the front-end, the symbolic pass, the code generator and the
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


def random_active_deck(seed: int):
    """Random decks with the device kinds random_deck() never emits: BJT (NPN / PNP), MOSFET (NMOS / PMOS, Level 1-3),
    coupled inductors (K over 2 or 3 windings) and magnetic-core inductors — on a resistive skeleton, one driven by a
    source, every node with a resistive path to ground.  The reference's BJT / PMOS models often overflow or fail to
    converge; those lanes are parity cases too (status, NaN / Inf classes, solve counts)."""
    rng = np.random.default_rng(10_000 + seed)
    flavour = ["bjt", "mos", "xfmr", "core"][seed % 4]
    lines = [f"* random active deck {seed} ({flavour})"]
    info = dict(flavour=flavour, has_inductor=flavour in ("xfmr",), nodes=0)
    r = lambda lo, hi: _eng(float(np.exp(rng.uniform(np.log(lo), np.log(hi)))), rng)
    if flavour == "bjt":
        pnp = bool(rng.integers(0, 2))
        vcc = float(rng.uniform(5, 12))
        lines.append(f"VCC 1 0 DC {vcc:.3f}")
        hi, lo = (f"{vcc:.3f}", f"{vcc - rng.uniform(0.55, 0.7):.3f}") if pnp else ("0", f"{rng.uniform(0.55, 0.7):.3f}")
        lines.append(f"VB 4 0 PULSE({hi} {lo} 2u 1u 1u 40u 100u)" if rng.random() < 0.7 else f"VB 4 0 DC {lo}")
        lines.append(f"RB 4 2 {r(5, 500)}")
        if pnp:
            lines += [f"RC 3 0 {r(200, 5e3)}", "Q1 3 2 1 QX"]
        else:
            lines += [f"RC 1 3 {r(200, 5e3)}", "Q1 3 2 0 QX"]
        lines.append(f".model QX {'PNP' if pnp else 'NPN'}(Is={rng.uniform(0.5, 5):.3f}e-14 Bf={int(rng.integers(50, 300))} "
                     f"Vaf={rng.uniform(40, 150):.1f} Ikf={rng.uniform(5, 50):.1f}m)")
        if rng.random() < 0.5:
            lines.append(f"RL 3 0 {r(1e3, 1e5)}")
        lines.append(".tran 1u 100u")
        info["nodes"] = 4
    elif flavour == "mos":
        level = int(rng.integers(1, 4))
        pmos = bool(rng.integers(0, 2))
        lines.append(f"VDD 1 0 DC {rng.uniform(3, 6):.3f}")
        src = rng.choice(["pulse", "sin"])
        lines.append("VG 2 0 PULSE(0 5 1u 200n 200n 4u 10u)" if src == "pulse" else f"VG 2 0 SIN(2.5 {rng.uniform(0.5, 2.4):.3f} 200k)")
        lines.append(f"RD 1 3 {r(80, 200) if pmos else r(1e3, 2e4)}")
        body = "0"
        if rng.random() < 0.4 and not pmos:
            lines += [f"VBB 5 0 DC -{rng.uniform(0.5, 3):.2f}", f"RS 4 0 {r(100, 2e3)}"]
            lines.append("M1 3 2 4 5 MX L=2u W=20u")
        else:
            lines.append(f"M1 3 2 0 {body} MX L={rng.choice(['1u', '2u', '5u'])} W={rng.choice(['10u', '20u', '40u'])}")
        extra = {1: "", 2: f" TOX=2e-8 UO={int(rng.integers(200, 700))} UCRIT=1e4 UEXP={rng.uniform(0.05, 0.2):.3f} VMAX={rng.uniform(3, 9):.1f}e4",
                 3: f" THETA={rng.uniform(0.01, 0.1):.3f} KAPPA={rng.uniform(0.1, 0.6):.2f} DELTA={rng.uniform(0, 1):.2f}"}[level]
        vto = -rng.uniform(0.5, 1.0) if pmos else rng.uniform(0.5, 1.0)
        lines.append(f".model MX {'PMOS' if pmos else 'NMOS'}(Level={level} VTO={vto:.3f} KP={rng.uniform(10, 60):.1f}u GAMMA={rng.uniform(0.2, 0.7):.2f} "
                     f"PHI={rng.uniform(0.55, 0.8):.2f} LAMBDA={rng.uniform(0, 0.05):.3f} CGSO=2e-10 CGDO=2e-10{extra})")
        lines.append(".tran 0.1u 10u")
        info["nodes"] = 5
    elif flavour == "xfmr":
        three = bool(rng.integers(0, 2))
        lines.append(f"Vin 1 0 SIN(0 {rng.uniform(2, 12):.2f} {rng.choice([500, 1000, 2000])})")
        lines += [f"Rp 1 2 {r(0.05, 2)}", f"Lp 2 0 {r(5e-2, 4e-1)}", f"Ls1 3 0 {r(1e-2, 1e-1)}", f"Rs1 3 4 {r(0.02, 1)}",
                  f"Rl1 4 0 {r(50, 2e4)}"]
        if three:
            lines += [f"Ls2 5 0 {r(1e-2, 1e-1)}", f"Rs2 5 6 {r(0.02, 1)}", f"Rl2 6 0 {r(50, 2e4)}"]
        lines.append(f"K1 Lp Ls1{' Ls2' if three else ''} {rng.uniform(0.8, 0.99):.3f}")
        lines.append(".tran 0.02m 0.6m")
        info["nodes"] = 6 if three else 4
    else:
        lines.append(f"Vin 1 0 SIN(0 {rng.uniform(2, 12):.2f} 1k)")
        lines += [f"Rp 1 2 {r(0.05, 2)}", f"Lp 2 0 core=CX turns={int(rng.integers(50, 400))}", f"Rs 3 4 {r(0.05, 1)}",
                  f"Ls 3 0 core=CX turns={int(rng.integers(50, 400))}", f"Rload 4 0 {r(100, 5e3)}",
                  f".model CX core(ms=1.6e6 a=1000 c=0.1 k=2000 area={rng.uniform(0.5, 3):.2f}e-4 len={rng.uniform(0.05, 0.3):.3f})",
                  f"K1 Lp Ls {rng.uniform(0.8, 0.99):.3f}", ".tran 10u 3m"]
        info["nodes"] = 4
    return "\n".join(lines) + "\n", info


def rc_ladder(sections: int) -> str:
    """Vin - (R - C to ground) x sections: n = sections + 2 unknowns, tridiagonal conductance block (no fill)."""
    lines = [f"* RC ladder, {sections} sections", "Vin 1 0 SIN(0 5 1k)"]
    for k in range(1, sections + 1):
        lines.append(f"R{k} {k} {k + 1} 100")
        lines.append(f"C{k} {k + 1} 0 100n")
    lines.append(".tran 0.01ms 3ms")
    return "\n".join(lines) + "\n"


def rlc_ladder(sections: int) -> str:
    """Vin - (R - L in series, C to ground) x sections: 2 nodes and one branch current per section."""
    lines = [f"* RLC ladder, {sections} sections", "Vin 1 0 SIN(0 5 1k)"]
    node = 1
    for k in range(1, sections + 1):
        lines.append(f"R{k} {node} {node + 1} 50")
        lines.append(f"L{k} {node + 1} {node + 2} 1m")
        lines.append(f"C{k} {node + 2} 0 100n")
        node += 2
    lines.append(".tran 0.01ms 1ms")
    return "\n".join(lines) + "\n"


def rc_mesh(rows: int, cols: int) -> str:
    """rows x cols grid of nodes, a resistor between neighbours, a capacitor to ground at every node, driven at one corner
    through a resistor: separators of a cut are whole rows / columns of the grid (several unknowns)."""
    def nd(i, j):
        return 2 + i * cols + j
    lines = [f"* RC mesh {rows} x {cols}", "Vin 1 0 PULSE(0 5 0.1m 0.05m 0.05m 0.8m 2m)", f"Rin 1 {nd(0, 0)} 100"]
    k = 0
    for i in range(rows):
        for j in range(cols):
            lines.append(f"C{i}_{j} {nd(i, j)} 0 {50 + 10 * ((i * 7 + j * 3) % 9)}n")
            if j + 1 < cols:
                k += 1; lines.append(f"Rh{k} {nd(i, j)} {nd(i, j + 1)} {100 + 20 * ((i + 2 * j) % 5)}")
            if i + 1 < rows:
                k += 1; lines.append(f"Rv{k} {nd(i, j)} {nd(i + 1, j)} {150 + 30 * ((2 * i + j) % 4)}")
    lines.append(f"Rload {nd(rows - 1, cols - 1)} 0 1k")
    lines.append(".tran 0.01ms 2ms")
    return "\n".join(lines) + "\n"


def random_linear_network(seed: int, nodes: int = 16):
    """A larger random LINEAR deck (for the cooperative mapping's partitioner): `nodes` nodes on a random spanning tree of
    resistors plus a few chords (loops make separators of several unknowns), a capacitor to ground at most nodes, a couple of
    inductors in series branches (branch unknowns), one or two sources of different waveforms.  Every node has a resistive
    path to ground through the tree and a load resistor."""
    rng = np.random.default_rng(50_000 + seed)
    lines = [f"* random linear network {seed}, {nodes} nodes"]
    src = ["SIN(0 5 1k)", "PULSE(0 5 0.1m 0.05m 0.05m 0.6m 1.5m)", "PWL(0 0 0.2m 0 0.5m 3.3 1m 3.3 1.2m -1 2m 0)"][seed % 3]
    lines.append(f"Vin 1 0 {src}")
    nr = nc = nl = 0
    nxt = nodes + 1                      # extra nodes for the series inductors
    n_ind = int(rng.integers(0, 3))
    ind_edges = set(rng.choice(np.arange(2, nodes + 1), size=n_ind, replace=False).tolist()) if n_ind else set()
    for k in range(2, nodes + 1):
        lo = max(1, k - 4)               # a "wide chain": parents are recent nodes, so the graph has small cuts
        other = int(rng.integers(lo, k))
        rv = float(np.exp(rng.uniform(np.log(50), np.log(5e3))))
        nr += 1
        if k in ind_edges:               # R in series with L through an extra node
            nl += 1
            lines.append(f"R{nr} {other} {nxt} {rv:.5g}")
            lines.append(f"L{nl} {nxt} {k} {float(np.exp(rng.uniform(np.log(2e-4), np.log(5e-3)))):.5g}")
            nxt += 1
        else:
            lines.append(f"R{nr} {other} {k} {rv:.5g}")
    for _ in range(max(1, nodes // 5)):  # chords
        a = int(rng.integers(2, nodes + 1))
        b = int(rng.integers(max(2, a - 5), min(nodes, a + 5) + 1))
        if a != b:
            nr += 1
            lines.append(f"R{nr} {a} {b} {float(np.exp(rng.uniform(np.log(200), np.log(2e4)))):.5g}")
    for k in range(2, nodes + 1):
        if rng.random() < 0.8:
            nc += 1
            lines.append(f"C{nc} {k} 0 {float(np.exp(rng.uniform(np.log(2e-8), np.log(5e-7)))):.5g}")
    nr += 1
    lines.append(f"R{nr} {nodes} 0 {float(np.exp(rng.uniform(np.log(500), np.log(5e3)))):.5g}")
    if seed % 2:
        lines.append(f"Iload {max(2, nodes // 2)} 0 SIN(0 1m 2k)")
    lines.append(".tran 0.01ms 2ms" if nl == 0 else ".tran 4e-6 4e-4")
    return "\n".join(lines) + "\n", dict(nodes=nodes, inductors=nl)


def diode_rc_ladder(sections: int) -> str:
    """Nonlinear larger-n deck: the RC ladder with a diode clamp to ground at every third node (and a reverse one at every
    fourth): Newton loops on a circuit of sections + 2 unknowns."""
    lines = [f"* diode-clamped RC ladder, {sections} sections", "Vin 1 0 SIN(0 5 1k)"]
    nd = 0
    for k in range(1, sections + 1):
        lines.append(f"R{k} {k} {k + 1} 100")
        lines.append(f"C{k} {k + 1} 0 100n")
        if k % 3 == 0:
            nd += 1
            lines.append(f"D{nd} {k + 1} 0 D")
        elif k % 4 == 0:
            nd += 1
            lines.append(f"D{nd} 0 {k + 1} D")
    lines.append(".tran 0.01ms 3ms")
    return "\n".join(lines) + "\n"


def mos_follower_chain(stages: int) -> str:
    """Nonlinear larger-n deck with a hub: `stages` NMOS source followers (Level 1) in a chain, all drains on the supply node
    — the supply node and its branch end up in the separator."""
    lines = [f"* NMOS source-follower chain, {stages} stages", "VDD 1 0 DC 5", "VG 2 0 PULSE(0 5 1u 100n 100n 5u 10u)"]
    prev = 2
    for k in range(1, stages + 1):
        out = 2 + k
        lines.append(f"M{k} 1 {prev} {out} 0 NMOS_T L=2u W=20u")
        lines.append(f"RS{k} {out} 0 10k")
        lines.append(f"CS{k} {out} 0 1p")
        prev = out
    lines.append(".model NMOS_T NMOS(Level=1 VTO=0.7 KP=20u LAMBDA=0.01)")
    lines.append(".tran 0.1u 10u")
    return "\n".join(lines) + "\n"
