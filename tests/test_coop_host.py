"""The cooperative mapping's generated code (codegen.cpp: emit_coop — one struct per sub-circuit, nested-dissection order)
compiled for the HOST and run part after part around an emulated exchange buffer: for arbitrary parameters, device state,
time and step the parts together must solve the system the reference-order elimination (Ckt::assemble_solve) solves, take
the same truncation-error decisions, leave the same device state behind and report the same result columns.  The
barrier / shared-memory choreography itself is what the GPU tests cover; everything arithmetic is covered here."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import parity_util as PU
from random_decks import diode_rc_ladder, mos_follower_chain, random_deck, random_linear_network, rc_ladder, rc_mesh, rlc_ladder

T = PU.T
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
#define __ldcs(p) (*(p))
struct double2 { double x, y; };
#include "MODELS"
struct TsbArgs { long long n_inst; const double* pv[128]; const double* U; double Uc[32]; double* coop_state; };
STRUCTS
template <class Part> static void run_a(Part& c, const TsbArgs& a, double time, double dt, double* xb, int* flags) {
    c.load(a, 0);
    bool gt, sm;
    c.lte_flags(dt, 1.0 / dt, 7.0, 0.07, gt, sm);
    c.eval_sources(time, 1.0);
    bool ok = c.phase_a(time, dt, 1.0 / dt, xb + Part::PART * TSB_COOP_NX * 32);
    flags[0] |= gt ? 1 : 0; flags[1] &= sm ? 1 : 0; flags[2] &= ok ? 1 : 0;
}
template <class Part> static void run_b(Part& c, const double* xb, double dt, double time, int* flags, double* xout, double* sout, double* row, const int* owner) {
    bool ok = c.phase_b(xb);
    flags[2] &= ok ? 1 : 0;
    c.load_state(dt); c.update_state();
    for (int u = 1; u <= Part::N; ++u) if (owner[u] == Part::PART || owner[u] < 0) {
        if (owner[u] < 0 && xout[u] == xout[u] && std::memcmp(&xout[u], &c.x[u], 8) != 0 && !(xout[u] == -12345.0)) flags[3] += 1;   // separator values: same bits in every part
        xout[u] = c.x[u];
    }
    for (int k = 0; k < NSTATE; ++k) if (SOWNER[k] == Part::PART) sout[k] = c.S[k];
    double r[Part::NOWN > 0 ? Part::NOWN : 1];
    c.signals(time + dt, r);
    for (int j = 0; j < Part::NOWN; ++j) row[c.col(j)] = r[j];
}
int main() {
    const int trials = 200;
    unsigned long long rng = 88172645463325252ULL;
    auto uni = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (double)(rng >> 11) / 9007199254740992.0; };
    const int owner[] = {OWNER};
    double worst = 0, worst_state = 0, worst_row = 0;
    int fails = 0, sep_mismatch = 0, flag_mismatch = 0;
    for (int t = 0; t < trials; ++t) {
        double U[NPAR + 1], V[128][1];
        TsbArgs a; a.n_inst = 1; a.U = U;
        const double nominal[] = {NOMINAL};
        for (int k = 0; k < NPAR; ++k) { U[k] = nominal[k]; if (k < 32) a.Uc[k] = U[k]; }
        for (int s = 0; s < NVAR; ++s) { const int k = VARIDX[s]; V[s][0] = nominal[k] * std::exp((uni() - 0.5) * 1.4); a.pv[s] = V[s]; }
        double state[NSTATE + Ckt::N + 2];
        double s2[NSTATE + 1];
        a.coop_state = state;
        Ckt c1;
        c1.load(a, 0); c1.init();
        for (int k = 0; k < NSTATE; ++k) { double v = (uni() - 0.5) * (NONLINEAR ? 1.2 : 6.0); c1.S[k] = v; state[k] = v; }
        for (int u = 0; u <= Ckt::N; ++u) { double v = u ? (uni() - 0.5) * 4.0 : 0.0; c1.x[u] = v; state[NSTATE + u] = v; }
        const double time = uni() * 2e-3, dt = std::exp(std::log(1e-9) + uni() * std::log(1e5));
        bool gt1, sm1;
        c1.lte_flags(dt, 1.0 / dt, 7.0, 0.07, gt1, sm1);
        c1.eval_sources(time, 1.0);
        bool ok1 = c1.assemble_solve<TSB_MODE_TRAN>(TSB_MODE_TRAN, time, dt, 1.0 / dt, 0.0);
        c1.load_state(dt); c1.update_state();
        double row1[TSB_COOP_NCOL];
        row1[0] = time + dt; c1.signals(row1 + 1);
        // the parts, one after the other around the exchange buffer
        static double xb[TSB_COOP_PARTS * TSB_COOP_NX * 32];
        int flags[4] = {0, 1, 1, 0};
        double x2[Ckt::N + 1], row2[TSB_COOP_NCOL];
        for (int u = 0; u <= Ckt::N; ++u) x2[u] = -12345.0;
        PARTS_DECL
        PARTS_A
        PARTS_B
        sep_mismatch += flags[3];
        if ((flags[0] != 0) != gt1 || (flags[1] != 0) != sm1) ++flag_mismatch;
        if (!ok1 || !flags[2]) { if (ok1 != (flags[2] != 0)) ++fails; continue; }
        double scale = 0;
        for (int i = 1; i <= Ckt::N; ++i) scale = std::fmax(scale, std::fabs(c1.x[i]));
        for (int i = 1; i <= Ckt::N; ++i) worst = std::fmax(worst, std::fabs(c1.x[i] - x2[i]) / (std::fabs(c1.x[i]) + 1e-6 * scale + 1e-300));
        double sscale = 0;
        for (int k = 0; k < NSTATE; ++k) sscale = std::fmax(sscale, std::fabs(c1.S[k]));
        for (int k = 0; k < NSTATE; ++k) worst_state = std::fmax(worst_state, std::fabs(c1.S[k] - s2[k]) / (std::fabs(c1.S[k]) + 1e-6 * sscale + 1e-300));
        double rscale = 0;
        for (int j = 0; j < TSB_COOP_NCOL; ++j) rscale = std::fmax(rscale, std::fabs(row1[j]));
        for (int j = 0; j < TSB_COOP_NCOL; ++j) worst_row = std::fmax(worst_row, std::fabs(row1[j] - row2[j]) / (std::fabs(row1[j]) + 1e-6 * rscale + 1e-300));
    }
    printf("worst %.3e state %.3e row %.3e fails %d sep_mismatch %d flag_mismatch %d\n", worst, worst_state, worst_row, fails, sep_mismatch, flag_mismatch);
    return 0;
}
'''


def coop_host_check(text, parts, tmp):
    ckt = T.Circuit.from_netlist(text)
    owner = ckt.coop_info(parts)
    if owner is None:
        return None
    b = ckt.batch(2)
    ov = PU.draws("x", ckt, 2, seed=5)
    for (d, p), v in ov.items():
        b.set_param(d, p, v)
    src = b.kernel_source(T.default_opts(strict_fp=0, min_blocks=2, coop_parts=parts))
    assert "tsb_coop_tran(TsbArgs a)" in src
    struct = re.search(r"struct Ckt \{.*?\n\};\n", src, re.S).group(0)
    defines = "\n".join(re.findall(r"^#define TSB_COOP_(?:PARTS|NX|NOWN_MAX|NCOL) .*$", src, re.M))
    partsrc = "\n".join(re.search(r"struct CoopPart%d \{.*?\n\};\n" % p, src, re.S).group(0) for p in range(parts))
    devs = ckt.devices()
    nominal = [v for d in devs for v in d["p"]]
    slots = [int(x) for x in re.findall(r"P\[(\d+)\] = __ldcs\(a\.pv\[\d+\]", struct)]
    nstate = max([int(k) + 1 for k in re.findall(r"o\[\(long long\)\d+ \* n_inst \+ inst\] = S\[(\d+)\]", struct)] or [0])     # real slots (S[] is never empty)
    # which part owns state slot k: from the parts' load() code
    sowner = [-1] * max(1, nstate)
    for p in range(parts):
        body = re.search(r"struct CoopPart%d \{.*?\n\};\n" % p, src, re.S).group(0)
        for k in re.findall(r"S\[(\d+)\] = a\.coop_state", body):
            sowner[int(k)] = p
    code = (HARNESS.replace("MODELS", os.path.join(ROOT, "toy-spice_b200", "csrc", "device", "models.cuh"))
            .replace("STRUCTS", defines + "\n#define TSB_INF INFINITY\n" + struct + partsrc if "TSB_INF" not in open(os.path.join(ROOT, "toy-spice_b200", "csrc", "device", "models.cuh")).read() else defines + "\n" + struct + partsrc)
            .replace("SOWNER", "((const int[]){" + ", ".join(map(str, sowner)) + "})")
            .replace("OWNER", ", ".join(map(str, owner)))
            .replace("NPAR", str(len(nominal))).replace("NOMINAL", ", ".join(repr(float(v)) for v in nominal) or "0")
            .replace("NVAR", str(len(slots))).replace("VARIDX", "((const int[]){" + ", ".join(map(str, slots or [0])) + "})")
            .replace("NSTATE", str(nstate)).replace("NONLINEAR", "1" if "HAS_NL = true" in struct else "0")
            .replace("PARTS_DECL", " ".join(f"CoopPart{p} q{p};" for p in range(parts)))
            .replace("PARTS_A", " ".join(f"run_a(q{p}, a, time, dt, xb, flags);" for p in range(parts)))
            .replace("PARTS_B", " ".join(f"run_b(q{p}, xb, dt, time, flags, x2, s2, row2, owner);" for p in range(parts))))
    cu = os.path.join(tmp, "h.cpp")
    open(cu, "w").write(code)
    exe = os.path.join(tmp, "h")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-DTSB_FAST_DIV", "-w", "-o", exe, cu], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True).stdout
    m = re.match(r"worst (\S+) state (\S+) row (\S+) fails (\d+) sep_mismatch (\d+) flag_mismatch (\d+)", out)
    assert m, out
    return tuple(float(x) for x in m.groups()[:3]) + tuple(int(x) for x in m.groups()[3:])


DECKS = {n: T.BUNDLED[n] for n in ("rl", "rlc")}
for _k in (4, 8, 12, 24):
    DECKS[f"ladder{_k}"] = rc_ladder(_k)
for _k in (2, 4, 7):
    DECKS[f"rlcladder{_k}"] = rlc_ladder(_k)
DECKS["mesh3x4"] = rc_mesh(3, 4)
DECKS["mesh4x5"] = rc_mesh(4, 5)
DECKS["random0"] = random_deck(0)[0]
for _s in range(8):
    DECKS[f"net{_s}"] = random_linear_network(_s, 12 if _s % 2 else 20)[0]
# circuits with nonlinear devices (the Newton-loop form of the cooperative mapping)
DECKS["diodeladder12"] = diode_rc_ladder(12)
DECKS["diodeladder24"] = diode_rc_ladder(24)
DECKS["moschain6"] = mos_follower_chain(6)
DECKS["moschain12"] = mos_follower_chain(12)
DECKS["mosfet1"] = T.BUNDLED["mosfet1"]


@pytest.mark.parametrize("parts", [2, 4, 8])
@pytest.mark.parametrize("name", sorted(DECKS))
def test_cooperative_solve_equals_reference_order(built, name, parts):
    with tempfile.TemporaryDirectory() as tmp:
        res = coop_host_check(DECKS[name], parts, tmp)
    if res is None:
        pytest.skip(f"no partition into {parts} sub-circuits for this netlist")
    worst, wstate, wrow, fails, sep_mismatch, flag_mismatch = res
    print(name, parts, res)
    assert fails == 0 and sep_mismatch == 0 and flag_mismatch == 0, res
    # two elimination orders of the same well-posed system: agreement to rounding x condition number
    assert worst < 1e-7 and wstate < 1e-7 and wrow < 1e-7, res


def test_partition_is_a_partition(built):
    """Every unknown has one owner or is in the separator; devices never couple two interiors (checked through the stamped
    pattern: an entry between two different interiors would make the plan refuse the partition)."""
    for name, text in DECKS.items():
        ckt = T.Circuit.from_netlist(text)
        for parts in (2, 4, 8):
            owner = ckt.coop_info(parts)
            if owner is None:
                continue
            assert len(owner) == ckt.n + 1
            ints = [sum(1 for u in owner[1:] if u == p) for p in range(parts)]
            assert all(c > 0 for c in ints), (name, parts, owner)
            assert all(-1 <= u < parts for u in owner[1:])
    # the 24-section ladder is a chain: one separator node per cut
    owner = T.Circuit.from_netlist(rc_ladder(24)).coop_info(4)
    assert sum(1 for u in owner[1:] if u < 0) == 3
