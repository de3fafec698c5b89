"""Decks for the AC analysis tests (ac.go).  The reference bundles one AC deck (bjt3.cir), which has a BJT: what its AC
analysis computes depends on the un-vendored sparse module's vector layout (include/tspice_b200.h: tsb_run_ac), so the
parity decks are linear networks: what every device's Stamp does in Mode == ACAnalysis, the frequency axes, the read-out."""

AC_DECKS = {
    # first-order low pass, decade sweep
    "ac_lowpass": """RC low pass
V1 1 0 AC 1 0
R1 1 2 1k
C1 2 0 1u
.ac DEC 31 1 1meg
""",
    # two sections, a voltage source with phase, a current source with phase, octave sweep
    "ac_ladder": """RC ladder, two sources
V1 1 0 AC 2 30
R1 1 2 1k
C1 2 0 1u
R2 2 3 2.2k
C2 3 0 470n
I1 0 3 AC 1m 45
R3 3 0 10k
.ac OCT 40 10 1meg
""",
    # twin-T notch: floating capacitors and resistors, linear sweep through the notch (f0 = 1/(2 pi R C) = 1591.5 Hz)
    "ac_twint": """Twin-T notch
Vin 1 0 AC 1 0
R1 1 2 10k
R2 2 4 10k
C3 2 0 20n
C1 1 3 10n
C2 3 4 10n
R3 3 0 5k
Rl 4 0 1meg
.ac LIN 61 1000 2200
""",
    # a DC source beside the AC one: its AC magnitude is zero, only the incidence entries are stamped
    "ac_two_v": """Divider between an AC and a DC source
V1 1 0 AC 1 90
V2 3 0 DC 5
R1 1 2 1k
R2 2 3 3k
C1 2 0 100n
.ac DEC 13 100 100k
""",
}

# the reference fails on these at the first frequency ("matrix solve error at f=..."): empty branch rows
AC_SINGULAR = {
    "ac_rl": """RL: the inductor's AC stamp is an admittance between its nodes, its branch row stays empty
V1 1 0 AC 1 0
R1 1 2 1k
L1 2 0 1m
.ac DEC 5 10 10k
""",
    "ac_core": """Core inductors and their coupling stamp nothing in AC mode
.model core1 CORE (ms=1.6e6 alpha=1e-3 a=1000 c=0.1 k=2000 area=1e-4 len=0.1)
Vin 1 0 AC 1 0
Rp 1 2 1
Lp 2 0 core=core1 turns=300
Ls 3 0 core=core1 turns=150
Rl 3 0 100
K1 Lp Ls 0.99
.ac DEC 5 10 10k
""",
}


def ac_draws(devices, n, seed=21):
    """Per-instance parameters of the GPU parity runs (and of the kernels __graft_entry__.build() pre-compiles for them): the
    SURVEY §8(d) draws of every R / C / L, plus magnitude and phase of the first AC source."""
    import importlib
    import numpy as np
    W = importlib.import_module("toy-spice_b200.workloads")
    ov = W.sweep_draws(devices, n, seed)
    rng = np.random.default_rng(5)
    src = [d["name"] for d in devices if d["kind"] in (3, 4) and len(d["p"]) >= 3]
    if src:
        ov[(src[0], 1)] = rng.uniform(0.5, 2.0, n)
        ov[(src[0], 2)] = rng.uniform(-180.0, 180.0, n)
    return ov
