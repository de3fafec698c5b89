"""Small decks written for the parity tests: the device-model code paths no bundled reference deck reaches.

  MOSFET Level 2 and Level 3 (mosfet.go:378-459), their finite-difference conductances (mosfet.go:505-533), PMOS
  (sign handling in calculateCurrents / calculateConductances / UpdateVoltages), body bias (vbs < 0: calculateVth's
  GAMMA term and the stale-gm gmbs of SURVEY Q14), PNP (bjt.go:220,283-313), the bundled bjt1 / bjt3 decks with the
  `.tran 1u 150u` card SURVEY §8(d)(4) supplies for BASELINE configs[3], and two-source decks for the nested DC sweep
  (dc.go:205-288).  Each deck is run as OP, DC sweep and transient by tests/test_gpu_parity.py and pinned by the
  independent NumPy restatement in tests/test_oracle.py.

EXTRA[name] = (netlist text, dict of analyses): "tran" -> None (the deck's own .tran card), "op" -> True,
"dc" -> (source, start, stop, inc) [on "dc_text" when present: the deck with that source as a DC source],
"dc2" -> ((outer source, start, stop, inc), (inner source, start, stop, inc))."""

from importlib import import_module

_B = import_module("toy-spice_b200.workloads").BUNDLED


def _with_card(text: str, card: str) -> str:
    """The bundled deck with its analysis card replaced ('bjt1/3 with a supplied .tran 1u 150u')."""
    lines = [ln for ln in text.splitlines() if not ln.strip().lower().startswith((".op", ".ac", ".tran", ".dc"))]
    return "\n".join(lines + [card]) + "\n"


def _dc_variant(text: str, src: str) -> str:
    """The deck with source `src` turned into a DC source: DCSweep's SetValue only reaches dcValue (vsource.go:241-244),
    which a PULSE source never reads — sweeping a pulsed source is a no-op in the reference."""
    out = []
    for ln in text.splitlines():
        f = ln.split()
        out.append(f"{f[0]} {f[1]} {f[2]} DC 0" if f and f[0].lower() == src.lower() else ln)
    return "\n".join(out) + "\n"


_PULSE = "PULSE(0 5 1u 200n 200n 4u 10u)"
_MOS = {
    # Level 2: mobility degradation (UCRIT / UEXP), velocity saturation (VMAX), forward-difference gm / gds / gmbs
    "mos2n": ("RD 1 3 10k", "N2 L=2u W=20u",
              "N2 NMOS(Level=2 VTO=0.8 KP=30u GAMMA=0.4 PHI=0.65 LAMBDA=0.02 TOX=2e-8 UO=600 UCRIT=1e4 UEXP=0.1 VMAX=5e4)"),
    # Level 3: THETA (mobility modulation), KAPPA (saturation field), DELTA (narrow width)
    "mos3n": ("RD 1 3 10k", "N3 L=2u W=20u",
              "N3 NMOS(Level=3 VTO=0.7 KP=25u GAMMA=0.5 PHI=0.6 LAMBDA=0.01 THETA=0.05 KAPPA=0.5 DELTA=0.8)"),
    # PMOS.  The reference flips the terminal voltages twice (UpdateVoltages multiplies by the type value, calculateCurrents
    # negates again, mosfet.go:640-665, 321-355), so its PMOS conducts for vg > vs with a NEGATIVE drain current; wired like
    # an NMOS with a small drain resistor the Level 1 model converges (V(3) rises above the rail) ...
    "mos1p": ("RD 1 3 100", "P1 L=2u W=40u", "P1 PMOS(Level=1 VTO=-0.8 KP=10u LAMBDA=0.02)"),
    # ... while the Level 2 / 3 PMOS Jacobian (forward differences with a double sign flip, mosfet.go:505-533) does not:
    # these two decks pin the FAILURE semantics (DC: status 3 at the second point; transient: step halving down to
    # minStep, then "failed to converge at t=..."), solve counts included.
    "mos2p": ("RD 1 3 100", "P2 L=2u W=40u", "P2 PMOS(Level=2 VTO=-0.8 KP=10u LAMBDA=0.02 TOX=2e-8 UO=250 UCRIT=1e4 UEXP=0.1 VMAX=4e4)"),
    "mos3p": ("RD 1 3 100", "P3 L=2u W=40u", "P3 PMOS(Level=3 VTO=-0.8 KP=10u LAMBDA=0.02 THETA=0.08 KAPPA=0.3)"),
}

EXTRA = {}
for _name, (_rd, _inst, _model) in _MOS.items():
    _t = f"* {_name}\nVDD 1 0 DC 5\nVG 2 0 {_PULSE}\n{_rd}\nM1 3 2 0 0 {_inst}\n.model {_model}\n.tran 0.1u 10u\n"
    EXTRA[_name] = (_t, dict(tran=None, op=True, dc=("VG", 0.0, 5.0, 0.25), dc_text=_dc_variant(_t, "VG")))

EXTRA.update({
    # NMOS level 2, source follower with the bulk on a negative rail: vbs < 0 (body effect in calculateVth, gmbs from the
    # stale gm of SURVEY Q14), overlap + bulk capacitances in the transient stamps
    "mos2body": ("* NMOS level 2 source follower with body bias\nVDD 1 0 DC 5\nVG 2 0 SIN(2.5 1.5 200k)\nVBB 5 0 DC -2\nRS 4 0 4.7k\n"
                 "M1 1 2 4 5 NB L=1u W=10u\n"
                 ".model NB NMOS(Level=2 VTO=0.6 KP=50u GAMMA=0.6 PHI=0.7 LAMBDA=0.03 TOX=1.5e-8 UO=500 UCRIT=2e4 UEXP=0.15 VMAX=8e4"
                 " CGSO=2e-10 CGDO=2e-10 CBD=5f CBS=5f)\n"
                 ".tran 0.05u 10u\n", dict(tran=None, dc=("VG", 0.0, 5.0, 0.25), op=True)),
    # NMOS level 1 with Meyer + junction capacitances derived from CJ / CJSW and the areas / perimeters (mosfet.go:553-565)
    "mos1caps": ("* NMOS level 1 with Meyer + junction capacitances\nVDD 1 0 DC 3.3\nVG 2 0 PULSE(0 3.3 0.5u 100n 100n 2u 5u)\nRD 1 3 4.7k\n"
                 "M1 3 2 0 0 NC L=1u W=10u AD=2e-11 AS=2e-11 PD=1.4e-5 PS=1.4e-5\n"
                 ".model NC NMOS(Level=1 VTO=0.6 KP=60u GAMMA=0.45 PHI=0.7 LAMBDA=0.04 TOX=2e-8 CGSO=3e-10 CGDO=3e-10 CGBO=1e-10"
                 " CJ=3e-4 CJSW=2e-10 MJ=0.5 PB=0.8)\n"
                 ".tran 0.05u 5u\n", dict(tran=None, op=True)),
    # PNP (bjt.go:220 sign, UpdateVoltages :295-299).  The reference BJT has no junction limiting:
    # pnp_op converges to a finite point from the base source; its DC sweep overflows to Inf / NaN (the Inf/NaN classes
    # must match point for point); pnp_tran NaNs after the base edge; pnp_small fails to converge at minStep.
    "pnp_op": ("* PNP operating point\nVCC 1 0 DC 10\nVB 2 0 DC 9.35\nRC 3 0 1k\nQ1 3 2 1 QP\n.model QP PNP(Is=1e-14 Bf=100 Vaf=80)\n.op\n",
               dict(op=True, dc=("VB", 9.0, 10.0, 0.05))),
    "pnp_tran": ("* PNP with base switching\nVCC 1 0 DC 10\nVB 4 0 PULSE(10 5 0 1u 1u 100u 200u)\nRB 4 2 10\nQ1 3 2 1 QP\nRC 3 0 10k\n"
                 ".model QP PNP(Is=1e-14 Bf=100 Vaf=100)\n.tran 1u 150u\n", dict(tran=None, op=True)),
    "pnp_small": ("* PNP with a small base drive\nVCC 1 0 DC 10\nVB 4 0 PULSE(10 9.35 0 1u 1u 100u 200u)\nRB 4 2 10\nQ1 3 2 1 QP\nRC 3 0 1k\n"
                  ".model QP PNP(Is=1e-14 Bf=100 Vaf=100)\n.tran 1u 150u\n", dict(tran=None)),
    # BASELINE configs[3]: bjt1 / bjt3 with the supplied transient card
    "bjt1_tran": (_with_card(_B["bjt1"], ".tran 1u 150u"), dict(tran=None)),
    "bjt3_tran": (_with_card(_B["bjt3"], ".tran 1u 150u"), dict(tran=None, op=True)),
    # two-source decks for the nested DC sweep (dc.go:205-288)
    "dio2src": ("* two-source diode deck\nV1 1 0 DC 0\nV2 3 0 DC 0\nR1 1 2 1k\nD1 2 3 D\nR2 2 0 10k\n.dc V1 0 2 0.25\n",
                dict(dc=("V1", 0.0, 2.0, 0.25), dc2=(("V1", 0.0, 2.0, 0.25), ("V2", -0.5, 0.5, 0.25)), op=True)),
    "mos_family": ("* NMOS output characteristics (VDS sweep for several VGS)\nVDS 1 0 DC 0\nVGS 2 0 DC 0\nM1 1 2 0 0 NF L=2u W=20u\n"
                   ".model NF NMOS(Level=1 VTO=0.7 KP=20u LAMBDA=0.02)\n.dc VDS 0 5 0.5\n",
                   dict(dc=("VDS", 0.0, 5.0, 0.5), dc2=(("VGS", 0.0, 4.0, 1.0), ("VDS", 0.0, 5.0, 0.5)))),
})
