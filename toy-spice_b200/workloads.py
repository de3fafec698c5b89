"""Workload decks: the netlists BASELINE.json's configs name (reference circuits/*.cir).

These are INPUT DATA (SPICE decks, 3-21 lines each), reproduced byte-for-byte because the
parity tests, bench.py and smoke() must run on GPU boxes where /root/reference does not exist.
No reference source code is copied.  Keys are the reference file names without ".cir".
"""

BUNDLED = {
    'bjt1': '* BJT DC Operating Point Test\nVCC 1 0 DC 10\nRB 1 2 10k\nRC 1 3 1k\nQ1 3 2 0 Q2N3904\n.model Q2N3904 NPN(Is=7.734e-15 Bf=416.4 Vaf=74.03 Cje=4.493p Cjc=3.638p Tf=0.1n Tr=10n)\n.op',
    'bjt2': '* BJT Test Circuit with Base Switching\nVCC 1 0 DC 10\nVB 4 0 PULSE(0 5 0 1u 1u 100u 200u)\nRB 4 2 10\nQ1 3 2 0 Q2N3904\nRC 1 3 10k\n\n.model Q2N3904 NPN(Is=1e-14 Bf=100 Vaf=100 Cje=10p Cjc=5p Tf=0.3n)\n.tran 1u 150u',
    'bjt3': '* Simple BJT AC Test Circuit\nVCC 1 0 DC 10\nVAC 2 0 AC 0.01\nRB 1 2 100k\nRC 1 3 10k\nQ1 3 2 0 Q2N3904\n.model Q2N3904 NPN(Is=1e-14 Bf=100 Vaf=100 Cje=8p Cjc=5p Tf=0.1n Tr=10n)\n.ac dec 10 10 1meg',
    'diode1': '* Diode Test Circuit\n.op\nvin 1 0 DC 5\nr1 1 2 1k\nd1 2 0 D\n',
    'diode2': '* Diode Test Circuit. Half wave rectifier\n.tran 0.1ms 3ms\nvin 1 0 sin(0 5 1k)\nd1 1 2 D\nr1 2 0 1k\n',
    'diode3': '* Diode Test Circuit. Half wave rectifier\n.dc Vin -1 3 0.1\nVin 1 0 DC 0\nd1 1 2 D\nr1 2 0 1k\n',
    'diode4': '* Diode Reverse Recovery Test Circuit\n.model MY_D D Tt=5n\n\n* 1V/-1V, td=20ns, tr=tf=1ns, PWdth=20ns, Period=40ns\nvpulse 1 0 pulse(1 -1 20ns 1ns 1ns 20ns 40ns)\nd1 1 2 MY_D\nr1 2 0 50\n\n*.op\n.tran 1ns 100ns\n',
    'diode5': '* Diode model params Test Circuit\n.model D1N4148 D (Is=4.352e-9 N=1.906 Rs=0.6458 Cj0=7.048e-13\n+ M=0.3333 Vj=0.869 Fc=0.5 Isr=3.333e-9 Nr=2)\nV1 anode 0 DC 5\nR1 anode n1 1k\nD1 n1 0 D1N4148\n.op\n',
    'idc': '* idc test\nI1 n1 0 DC 1m\nR1 n1 0 1k\n.op',
    'ipulse': '* ipulse test\nIpulse n1 0 PULSE(0 5m 2m 0.5m 0.5m 5m 10m)\nR1 n1 0 1k\n.tran 0.1m 30m',
    'ipwl': '* ipwl test circuit\nIpwl n1 0 PWL(0 0 2m 0 2.5m 3.3m 5m 3.3m 5.5m 0 10m 0)\nR1 n1 0 1k\n.tran 0.1m 15m',
    'isin': '* isin test circuit\nIsin n1 0 SIN(0 2m 1k 0)  ; offset=0, amplitude=2mA, freq=1kHz, phase=0\nR1 n1 0 1k\n.tran 0.1ms 3ms',
    'mosfet1': '* Simple NMOS Test Circuit\n\nVDD 1 0 DC 5\nVG 2 0 PULSE(0 5 1u 100n 100n 5u 10u)\n\nRD 1 3 10k\nM1 3 2 0 0 NMOS_Test L=2u W=20u\n\n.model NMOS_Test NMOS(Level=1 VTO=0.7 KP=20u LAMBDA=0.01)\n\n.tran 0.1u 10u\n',
    'rc': '* RC Test\n.tran 0.01ms 3ms\nvin 1 0 sin (0 5 1k)\n*.op\n*vin 1 0 DC 5\n*vin 1 0 ac 1\n*.ac dec 10 1 1meg\n*.ac lin 100 1 1meg\nr1 1 2 100\nc1 2 0 1u',
    'rl': '* RL Test\n.tran 10u 2ms\nVin 1 0 SIN (0 5 1k)\nR1 1 2 100\nL1 2 0 1m\n',
    'rlc': '* RLC Test\n.tran 0.01m 2ms\nVin 1 0 SIN(0 5 1k)\nR1 1 2 100\nL1 2 3 1m\nC1 3 0 1u\n',
    'rr': '* RR Test\n.tran 0.1m 3ms\n*Vin 1 0 SIN(0 5 1k)\n*.op\nVin 1 0 DC 5\nR1 1 2 1k\nR2 2 0 1k',
    'transformer1': '* Transformer Test Circuit with 2:1 ratio\nVin 1 0 sin(0 10 1k)\n\nRp_leak 1 2 0.1\nLp 2 0 200m\n\nLs 3 0 50m\nRs_leak 3 4 0.05\n\nRload 4 0 10k\n\nK1 Lp Ls 0.95\n\n.tran 0.01m 3m',
    'transformer2': '* Transformer Test Circuit with 2:1 ratio\nVin 1 0 sin(0 10 1k)\n\nRp_leak 1 2 0.1\nLp 2 0 200m\n\nLs1 3 0 50m\nRs1_leak 3 4 0.05\nRload1 4 0 100\n\nLs2 5 0 50m\nRs2_leak 5 6 0.05\nRload2 6 0 100\n\nK1 Lp Ls1 Ls2 0.95\n\n.tran 10u 3m',
    'transformer3': '* Nonlinear Transformer Test Circuit with 2:1 ratio\nVin 1 0 sin(0 10 1k)\n\nRp_leak 1 2 0.1\nLp 2 0 core=CORE1 turns=300\n\nRs_leak 3 4 0.1\nLs 3 0 core=CORE1 turns=150\nRload 4 0 1000\n\n.model CORE1 core(\n+ ms=1.6e6\n+ alpha=1e-3\n+ a=1000\n+ c=0.1\n+ k=2000\n+ area=1e-4\n+ len=0.1)\n\nK1 Lp Ls 0.95\n\n.tran 10u 3m',
    'vpulse': '* vpulse test circuit\nVpulse n1 0 PULSE(0 5 2ms 0.5ms 0.5ms 5ms 10ms)\nR1 n1 0 1k\n\n.tran 0.1ms 30ms',
    'vpwl': '* vpwl test circuit\nVpwl n1 0 PWL(0 0 2ms 0 2.5ms 3.3 5ms 3.3 5.5ms 0 10ms 0)\nR1 n1 0 1k\n\n.tran 0.1ms 15ms',
}


# Synthetic sweep definitions of SURVEY.md §8(d): which parameters of which device kinds vary per
# instance, and how they are drawn around the netlist's nominal value.  Device kinds use the
# numbering of include/tspice_b200.h.  Entries: kind -> [(param index, distribution, a, b)], with
# "logu" = nominal * LogUniform[a, b], "u" = Uniform[a, b], "n" = Normal(a, b).
SWEEP_SPEC = {
    0: [(0, "logu", 0.5, 2.0)],                       # R
    1: [(0, "logu", 0.5, 2.0)],                       # C
    2: [(0, "logu", 0.5, 2.0)],                       # L
    8: [(0, "u", 0.8, 0.99)],                         # K coupling coefficient
    5: [(0, "logu", 0.1, 10.0), (1, "u", 1.0, 2.0)],  # D: Is, N
    6: [(0, "logu", 0.1, 10.0), (2, "u", 0.97, 0.995), (5, "u", 50.0, 150.0)],   # Q: Ies, alphaF, Vaf
    7: [(0, "n", 0.7, 0.05), (1, "logu", 0.8, 1.25), (4, "u", 0.0, 0.02)],       # M: VTO, KP, LAMBDA
    9: [(1, "logu", 0.5, 2.0)],                       # core area
}
SWEEP_SEEDS = {"rc": 1234, "rlc": 1234, "rl": 1234, "rr": 1234, "diode": 2345, "bjt": 3456, "mosfet": 3456,
               "transformer": 4567}


def sweep_seed(name: str) -> int:
    for k, v in SWEEP_SEEDS.items():
        if name.startswith(k):
            return v
    return 999


def sweep_draws(devices, n: int, seed: int):
    """{(device name, param index): float64[n]} in device order (PCG64, instance-major per parameter)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for d in devices:
        for (pi, dist, a, b) in SWEEP_SPEC.get(d["kind"], []):
            nominal = d["p"][pi]
            if dist == "logu":
                v = nominal * np.exp(rng.uniform(np.log(a), np.log(b), n))
            elif dist == "u":
                v = rng.uniform(a, b, n)
            else:
                v = rng.normal(a, b, n)
            out[(d["name"], pi)] = v
    return out


def rc_ladder(sections: int) -> str:
    """Synthetic larger-n deck (not a reference circuit): Vin - (R - C to ground) x sections, i.e.
    sections + 2 unknowns and 2 * sections + 3 result columns — past n ~ 14 more than one thread's registers hold."""
    lines = [f"* RC ladder, {sections} sections", "Vin 1 0 SIN(0 5 1k)"]
    for k in range(1, sections + 1):
        lines.append(f"R{k} {k} {k + 1} 100")
        lines.append(f"C{k} {k + 1} 0 100n")
    lines.append(".tran 0.01ms 3ms")
    return "\n".join(lines) + "\n"


def diode_rc_ladder(sections: int) -> str:
    """The same ladder with a diode clamp to ground at every third node and a reverse one at every fourth (not multiples of
    three): a larger-n deck with Newton loops."""
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
