"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.

netlist.py — Python restatement of the reference's host front-end:
  pkg/netlist/parser.go   Parse / parseLine / parseDotOperator / parseModel / parseElement /
                          parseVoltageSource / parseCurrentSource / ParseValue / CreateDevice
  pkg/circuit/circuit.go  AssignNodeBranchMaps (:48-71), SetupDevices device order (:78-152)

It produces the flat device table both the oracle core (tspice_oracle.cpp) and the product's
C-ABI (include/tspice_b200.h) consume.  It is deliberately independent from the product's C++
front-end (toy-spice_b200/csrc/host/netlist.cpp): tests compare the two tables on every
bundled netlist.  Quirks Q18-Q20 of SURVEY.md §3.6 are reproduced.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

# device kinds (shared numbering with include/tspice_b200.h)
K_R, K_C, K_L, K_V, K_I, K_D, K_Q, K_M, K_K, K_LCORE = range(10)
SRC_DC, SRC_SIN, SRC_PULSE, SRC_PWL = range(4)
AN_OP, AN_TRAN, AN_AC, AN_DC = range(4)

_UNIT = {"T": 1e12, "G": 1e9, "meg": 1e6, "K": 1e3, "k": 1e3, "m": 1e-3, "u": 1e-6,
         "n": 1e-9, "p": 1e-12, "f": 1e-15}                       # parser.go:62-73
_VALUE_RE = re.compile(r"^([-+]?\d*\.?\d+(?:[eE][-+]?\d+)?)(meg|[TGMKkmunpf])?s?$")   # parser.go:728


class NetlistError(ValueError):
    pass


def parse_value(val: str) -> float:
    """parser.go:727-748.  `M` matches the regexp but has no multiplier (Q18)."""
    m = _VALUE_RE.match(val.strip())
    if m is None:
        raise NetlistError(f"invalid value format: {val}")
    num = float(m.group(1))
    if m.group(2):
        mult = _UNIT.get(m.group(2))
        if mult is not None:
            num *= mult
    return num


@dataclass
class Element:
    type: str
    name: str
    nodes: list = field(default_factory=list)
    value: float = 0.0
    params: dict = field(default_factory=dict)


@dataclass
class Netlist:
    title: str = ""
    elements: list = field(default_factory=list)
    models: dict = field(default_factory=dict)      # name -> (type, params)
    analysis: int = AN_OP
    tran: dict = field(default_factory=lambda: dict(tstep=0.0, tstop=0.0, tstart=0.0, tmax=0.0, uic=False))
    dc: dict = field(default_factory=lambda: dict(source="", start=0.0, stop=0.0, inc=0.0))
    dc2: dict = field(default_factory=lambda: dict(source="", start=0.0, stop=0.0, inc=0.0))     # hardening: second .dc source
    ac: dict = field(default_factory=lambda: dict(sweep="", points=0, fstart=0.0, fstop=0.0))


def _fields(s: str) -> list:
    return s.split()


def parse(text: str) -> Netlist:
    """parser.go:75-158 (bufio.Scanner line splitting: '\\n' with an optional trailing '\\r')."""
    nl = Netlist()
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    lines = [ln[:-1] if ln.endswith("\r") else ln for ln in lines]
    if not lines:
        return nl
    title = lines[0]
    if title.startswith("*"):
        title = title[1:]
    nl.title = title.strip()

    current = ""
    continuation = False
    for raw in lines[1:]:
        line = raw.strip()
        if len(line) == 0:
            if current != "":
                _parse_line(nl, current)
                current = ""
                continuation = False
            continue
        idx = line.find("*")
        if idx >= 0:
            line = line[:idx].strip()
            if len(line) == 0:
                continue
        if line.startswith("*"):          # unreachable after the truncation above; kept for shape
            continue
        if line.startswith("+"):
            line = line[1:].strip()
            if current != "":
                current += " " + line
            continuation = True
            continue
        if continuation and raw.startswith(" "):
            line = line.strip()
            if current != "":
                current += " " + line
            continue
        if current != "":
            _parse_line(nl, current)
        current = line
        continuation = False
    if current != "":
        _parse_line(nl, current)
    return nl


def _parse_line(nl: Netlist, line: str) -> None:
    line = re.sub(r"\s+", " ", line)
    if getattr(nl, "_ended", False):
        return
    if line.strip().lower() == ".end":          # HARDENING (parser.go:155 "TODO: .END"): the deck ends here; the reference errors
        nl._ended = True
        return
    if line.startswith("."):
        _parse_dot(nl, line)
        return
    nl.elements.append(_parse_element(line))


def _parse_dot(nl: Netlist, line: str) -> None:
    f = _fields(line)
    if not f:
        raise NetlistError("invalid analysis command")
    cmd = f[0].lower()
    if cmd == ".model":
        _parse_model(nl, f[1:])
    elif cmd == ".op":
        nl.analysis = AN_OP
    elif cmd == ".tran":
        nl.analysis = AN_TRAN
        if len(f) < 3:
            raise NetlistError("insufficient tran parameters, need at least tstep and tstop")
        nl.tran["tstep"] = parse_value(f[1])
        nl.tran["tstop"] = parse_value(f[2])
        for i in range(3, len(f)):
            if f[i] == "uic":
                nl.tran["uic"] = True
                continue
            if i == 3:
                nl.tran["tstart"] = parse_value(f[i])
            if i == 4:
                nl.tran["tmax"] = parse_value(f[i])
        if nl.tran["tmax"] == 0:
            nl.tran["tmax"] = nl.tran["tstep"]
    elif cmd == ".ac":                       # parser.go:238-261
        nl.analysis = AN_AC
        if len(f) < 5:
            raise NetlistError("insufficient AC parameters")
        sweep = f[1].upper()
        if sweep not in ("DEC", "OCT", "LIN"):
            raise NetlistError(f"invalid sweep type: {sweep}")
        try:
            pts = int(f[2])
        except ValueError:
            raise NetlistError("invalid number of points")
        nl.ac = dict(sweep=sweep, points=pts, fstart=parse_value(f[3]), fstop=parse_value(f[4]))
    elif cmd == ".dc":
        nl.analysis = AN_DC
        if len(f) < 5:
            raise NetlistError("insufficient DC sweep parameters")
        nl.dc = dict(source=f[1], start=parse_value(f[2]), stop=parse_value(f[3]), inc=parse_value(f[4]))
        # HARDENING beyond the reference (its parser stops after the first source; cmd/spice/main.go:325 is
        # ready for DCParam.Source2): `.dc src1 start stop inc src2 start stop inc` -> nested sweep, src1 the outer loop
        if len(f) >= 9:
            nl.dc2 = dict(source=f[5], start=parse_value(f[6]), stop=parse_value(f[7]), inc=parse_value(f[8]))
    else:
        raise NetlistError(f"unsupported analysis type: {f[0]}")


_MODEL_DEFAULTS = {   # parser.go:347-431
    "D": dict(**{"is": 1e-14}, n=1.0, rs=0.0, cj0=0.0, m=0.5, vj=1.0, bv=100.0, eg=1.11, xti=3.0, tt=0.0, fc=0.5),
    "CORE": dict(ms=1.6e6, alpha=1e-3, a=1000.0, c=0.1, k=2000.0, tc=1043.0, beta=0.0, area=1e-4, len=0.1),
    "BJT": dict(**{"is": 1e-16}, bf=100.0, br=1.0, nf=1.0, nr=1.0, vaf=100.0, var=100.0, ikf=0.01, ikr=0.01,
                rc=0.0, re=0.0, rb=0.0, cje=0.0, vje=0.75, mje=0.33, cjc=0.0, vjc=0.75, mjc=0.33,
                tf=0.0, tr=0.0, xtb=0.0, eg=1.11, xti=3.0),
    "MOS": dict(level=1, vto=0.7, kp=2e-5, gamma=0.5, phi=0.6, **{"lambda": 0.01}, rd=0.0, rs=0.0, cbd=0.0,
                cbs=0.0, **{"is": 1e-14}, pb=0.8, cgso=0.0, cgdo=0.0, cgbo=0.0, cj=0.0, mj=0.5, cjsw=0.0,
                mjsw=0.33, tox=1e-7, l=10e-6, w=10e-6),
}


def _parse_model(nl: Netlist, f: list) -> None:
    """parser.go:285-451 (incl. the `D (` key bug, Q19)."""
    if len(f) < 2:
        raise NetlistError("insufficient model parameters")
    f = list(f)
    name = f[0]
    type_field = f[1]
    has_open = False
    if "(" in type_field:
        parts = type_field.split("(", 1)
        mtype = parts[0].upper()
        has_open = True
        if len(parts) > 1:
            f = f[:2] + [parts[1]] + f[2:]
    else:
        mtype = type_field.upper()
    if mtype not in ("D", "CORE", "NPN", "PNP", "NMOS", "PMOS"):
        raise NetlistError(f"unsupported model type: {mtype}")
    param_str = ""
    if has_open:
        pp = f[2:]
        if pp and pp[-1].endswith(")"):
            pp[-1] = pp[-1][:-1]
        param_str = " ".join(pp)
    elif len(f) > 2:
        param_str = " ".join(f[2:])
        if param_str.endswith(")"):
            param_str = param_str[:-1]
    param_str = re.sub(r"\*.*$", "", param_str).strip()
    if mtype == "D":
        params = dict(_MODEL_DEFAULTS["D"])
    elif mtype == "CORE":
        params = dict(_MODEL_DEFAULTS["CORE"])
    elif mtype in ("NPN", "PNP"):
        params = dict(_MODEL_DEFAULTS["BJT"])
        if mtype == "PNP":
            params["type"] = 1.0
    else:
        params = dict(_MODEL_DEFAULTS["MOS"])
        if mtype == "PMOS":
            params["type"] = 1.0
    for pair in param_str.split():
        parts = pair.split("=")
        if len(parts) != 2:
            continue
        params[parts[0].strip().lower()] = parse_value(parts[1].strip())
    nl.models[name] = (mtype, params)


def _parse_source(f: list, typ: str) -> Element:
    """parser.go:563-725 (voltage and current sources share the shape)."""
    if len(f) < 4:
        raise NetlistError("insufficient source parameters")
    e = Element(type=typ, name=f[0], nodes=[f[1], f[2]])
    remaining = " ".join(f[3:]).replace("(", " ( ").replace(")", " ) ")
    words = remaining.split()
    if not words:
        raise NetlistError("missing source type")
    kind = words[0].upper()
    if kind == "DC":
        if len(words) < 2:
            raise NetlistError("missing DC value")
        e.params["type"] = "dc"
        e.value = parse_value(words[1])
    elif kind in ("SIN", "PULSE", "PWL"):
        e.params["type"] = kind.lower()
        e.params[kind.lower()] = " ".join(words[1:]).strip("() ")
    elif kind == "AC":
        if len(words) < 2:
            raise NetlistError("missing AC magnitude")
        e.params["type"] = "ac"
        e.value = parse_value(words[1])
        e.params["phase"] = words[2] if len(words) > 2 else "0"
    else:
        raise NetlistError(f"unsupported source type: {words[0]}")
    return e


def _parse_element(line: str) -> Element:
    """parser.go:453-561."""
    f = _fields(line)
    if len(f) < 3:
        raise NetlistError(f"invalid element format: {line}")
    typ = f[0][0].upper()
    e = Element(type=typ, name=f[0])
    if typ == "V":
        return _parse_source(f, "V")
    if typ == "I":
        return _parse_source(f, "I")
    if typ == "L":
        e.nodes = f[1:3]
        for tok in f[3:]:
            pair = tok.split("=")
            if len(pair) == 2:
                e.params[pair[0].lower()] = pair[1]
            elif "=" not in tok:
                e.value = parse_value(tok)
        return e
    if typ == "K":
        if len(f) < 4:
            raise NetlistError("insufficient mutual coupling parameters")
        k = parse_value(f[-1])
        if k < -1 or k > 1:
            raise NetlistError("coupling coefficient must be between -1 and 1")
        names = f[1:-1]
        if len(names) < 2:
            raise NetlistError("mutual coupling requires at least two inductors")
        for i, nm in enumerate(names):
            e.params[f"ind{i + 1}"] = nm
        e.value = k
        return e
    if typ == "D":
        e.nodes = f[1:3]
        if len(f) > 3:
            e.params["model"] = f[3]
        for extra in f[4:]:                      # HARDENING: inline Is= / N= / Tt= overrides (the reference ignores these fields)
            kv = extra.split("=")
            if len(kv) == 2:
                e.params["inline_" + kv[0].lower()] = kv[1]
        return e
    if typ == "Q":
        if len(f) < 4:
            raise NetlistError("insufficient BJT parameters")
        e.nodes = f[1:4]
        if len(f) > 4:
            e.params["model"] = f[4]
        return e
    if typ == "M":
        if len(f) < 6:
            raise NetlistError("insufficient MOSFET parameters")
        e.nodes = f[1:5]
        e.params["model"] = f[5]
        for tok in f[6:]:
            parts = tok.split("=")
            if len(parts) == 2:
                e.params[parts[0].lower()] = parts[1]
        return e
    e.nodes = f[1:-1]
    e.value = parse_value(f[-1])
    return e


# ----------------------------------------------------------------------------- device table
@dataclass
class DeviceRow:
    kind: int
    name: str
    nodes: list          # node indices (0 = ground)
    branch: int
    p: list              # doubles, layout per kind (see include/tspice_b200.h)
    ip: list             # ints


@dataclass
class Plan:
    node_map: dict
    branch_map: dict
    devices: list        # DeviceRow, netlist order (K rows keep their netlist position; stamped last)
    netlist: Netlist

    @property
    def n_nodes(self):
        return len(self.node_map)

    @property
    def n_branches(self):
        return len(self.branch_map)


def _src_params(e: Element):
    t = e.params.get("type")
    if t == "dc":
        return SRC_DC, [e.value]
    if t == "sin":
        sp = e.params["sin"].split()
        if len(sp) < 3:
            raise NetlistError("insufficient SIN parameters")
        vals = [parse_value(sp[0]), parse_value(sp[1]), parse_value(sp[2])]
        vals.append(parse_value(sp[3]) if len(sp) > 3 else 0.0)
        return SRC_SIN, vals
    if t == "pulse":
        pp = e.params["pulse"].split()
        if len(pp) < 7:
            raise NetlistError("insufficient PULSE parameters")
        return SRC_PULSE, [parse_value(x) for x in pp[:7]]
    if t == "pwl":
        pw = e.params["pwl"].split()
        if len(pw) < 4 or len(pw) % 2 != 0:
            raise NetlistError("insufficient or invalid PWL parameters")
        vals = [parse_value(x) for x in pw]
        for i in range(2, len(vals), 2):
            if vals[i] <= vals[i - 2]:
                raise NetlistError("PWL time points must be strictly increasing")
        return SRC_PWL, vals
    if t == "ac":
        parse_value(e.params["phase"])
        # NewACVoltageSource(name, nodes, 0, mag, phase): DC 0 in OP / DC / tran; magnitude and phase (degrees) ride along
        return SRC_DC, [0.0, e.value, parse_value(e.params["phase"])]
    raise NetlistError(f"unsupported source type: {t}")


_MOS_KEYS = ["vto", "kp", "gamma", "phi", "lambda", "w", "l", "tox", "cgso", "cgdo", "cgbo", "cbd", "cbs", "cj",
             "cjsw", "as", "ad", "ps", "pd", "mj", "pb", "uo", "ucrit", "uexp", "vmax", "theta", "eta", "kappa",
             "delta"]
_MOS_DEV_DEFAULTS = dict(vto=0.7, kp=2e-5, gamma=0.5, phi=0.6, **{"lambda": 0.01}, w=10e-6, l=10e-6, tox=1e-7,
                         cgso=0.0, cgdo=0.0, cgbo=0.0, cbd=0.0, cbs=0.0, cj=0.0, cjsw=0.0, ad=0.0, ps=0.0,
                         pd=0.0, mj=0.5, pb=0.8, uo=600.0, ucrit=1e4, uexp=0.0, vmax=0.0, theta=0.0, eta=0.0,
                         kappa=0.2, delta=0.0, **{"as": 0.0})          # mosfet.go:144-208


def build_plan(nl: Netlist) -> Plan:
    """circuit.go:48-71 numbering + parser.go:752-915 CreateDevice parameter resolution."""
    node_map: dict = {}
    for e in nl.elements:
        for nm in e.nodes:
            if nm in ("0", "gnd"):
                continue
            if nm not in node_map:
                node_map[nm] = len(node_map) + 1
    branch_map: dict = {}
    b = len(node_map) + 1
    for e in nl.elements:
        if e.type in ("V", "L"):
            branch_map[e.name] = b
            b += 1

    rows = []
    index_of = {}
    for e in nl.elements:
        nodes = [0 if nm in ("0", "gnd") else node_map[nm] for nm in e.nodes]
        br = branch_map.get(e.name, 0)
        t = e.type
        if t == "R":
            row = DeviceRow(K_R, e.name, nodes, 0, [e.value], [])
        elif t == "C":
            row = DeviceRow(K_C, e.name, nodes, 0, [e.value], [])
        elif t == "L":
            if "core" in e.params:
                core = e.params["core"]
                if core not in nl.models:
                    raise NetlistError(f"undefined core model for inductor {e.name}: {core}")
                mtype, mp = nl.models[core]
                if mtype != "CORE":
                    raise NetlistError(f"invalid core model type for inductor {e.name}: {mtype}")
                turns = 100
                try:
                    turns = int(e.params.get("turns", "100"))     # strconv.Atoi
                except ValueError:
                    turns = 100
                area = mp.get("area", 1e-4)
                length = mp.get("len", 0.1)
                row = DeviceRow(K_LCORE, e.name, nodes, br, [float(turns), area, length], [])
            else:
                row = DeviceRow(K_L, e.name, nodes, br, [e.value], [])
        elif t == "K":
            row = DeviceRow(K_K, e.name, [], 0, [e.value], [])      # inductor indices resolved below
        elif t == "D":
            p = dict(**{"is": 1e-14}, n=1.0, tt=0.0)
            mname = e.params.get("model")
            if mname in nl.models:
                mp = nl.models[mname][1]
                for k in p:
                    if k in mp:
                        p[k] = mp[k]
            for k in p:                          # HARDENING (parser.go:530 "TODO: Inline parameters"): D1 a k MODEL Is=.. N=.. Tt=..
                if "inline_" + k in e.params:
                    p[k] = parse_value(e.params["inline_" + k])
            if len(e.nodes) != 2:
                raise NetlistError(f"diode {e.name}: requires exactly 2 nodes")
            row = DeviceRow(K_D, e.name, nodes, 0, [p["is"], p["n"], p["tt"]], [])
        elif t == "Q":
            p = dict(ies=1e-15, ics=1e-15, alphaf=0.98, ikf=1e-3, ikr=1e-3, vaf=50.0, var=50.0)   # bjt.go:87-108
            pnp = 0
            mname = e.params.get("model")
            if mname in nl.models:
                mp = nl.models[mname][1]
                for k in p:
                    if k in mp:
                        p[k] = mp[k]
                if "type" in mp and mp["type"] == 1.0:
                    pnp = 1
            row = DeviceRow(K_Q, e.name, nodes, 0,
                            [p["ies"], p["ics"], p["alphaf"], p["ikf"], p["ikr"], p["vaf"], p["var"], 1.0, 1.0], [pnp])
        elif t == "M":
            p = dict(_MOS_DEV_DEFAULTS)
            level, pmos = 1, 0
            mname = e.params.get("model")
            if mname is None:
                raise NetlistError(f"mosfet {e.name}: model not specified")
            if mname in nl.models:
                mp = nl.models[mname][1]
                if "level" in mp:
                    level = int(mp["level"])
                if "type" in mp:
                    pmos = 1 if mp["type"] == 1.0 else 0
                for k in p:
                    if k in mp:
                        p[k] = mp[k]
            for k in ("l", "w"):
                if k in e.params:
                    try:
                        p[k] = parse_value(e.params[k])
                    except NetlistError:
                        pass
            row = DeviceRow(K_M, e.name, nodes, 0, [p[k] for k in _MOS_KEYS], [level, pmos])
        elif t in ("V", "I"):
            st, vals = _src_params(e)
            row = DeviceRow(K_V if t == "V" else K_I, e.name, nodes, br, vals, [st])
        else:
            raise NetlistError(f"unsupported device type: {t}")
        index_of[e.name] = len(rows)
        rows.append(row)

    for e, row in zip(nl.elements, rows):
        if row.kind != K_K:
            continue
        names = []
        i = 1
        while f"ind{i}" in e.params:
            names.append(e.params[f"ind{i}"])
            i += 1
        for nm in names:
            if nm not in index_of or rows[index_of[nm]].kind not in (K_L, K_LCORE):
                raise NetlistError(f"inductor {nm} not found for mutual coupling {e.name}")
            row.ip.append(index_of[nm])
    return Plan(node_map, branch_map, rows, nl)


def signal_names(plan: Plan, analysis: int) -> list:
    """Canonical column order used by the oracle core and the product (the reference returns
    a map, circuit.go:242-273 / op.go:235-248; only key -> series matters)."""
    by_idx = sorted(plan.node_map.items(), key=lambda kv: kv[1])
    if analysis == AN_AC:        # StoreACResult (anlysis.go:87-111): <key>_MAG, <key>_PHASE for node voltages and V-source currents
        keys = [f"V({k})" for k, _ in by_idx] + [f"I({r.name})" for r in plan.devices if r.kind == K_V]
        return ["FREQ"] + [k + sfx for k in keys for sfx in ("_MAG", "_PHASE")]
    names = [f"V({k})" for k, _ in by_idx]
    names += [f"I({k})" for k, _ in sorted(plan.branch_map.items(), key=lambda kv: kv[1])]
    if analysis == AN_OP:
        return names
    names += [f"I({r.name})" for r in plan.devices if r.kind == K_R]
    return (["TIME"] if analysis == AN_TRAN else ["SWEEP1"]) + names
