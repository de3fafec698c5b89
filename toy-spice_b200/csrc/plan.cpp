// plan.cpp — freezes a numbered device table into everything the GPU kernels are specialised on.
//
//   stamp order            circuit.go:83-152 (netlist order, mutual couplings last)
//   stamp call sequence    the ordered AddElement/AddRHS calls of every device Stamp() (pkg/device/*.go)
//   Translate numbering    first-touch order of the setup stamp + SetupElements (circuit.go:154-160,
//                          matrix/circuit.go:57-63) — what Sparse's ext->int map looks like
//   pivot order            the reference's first Factor() happens in iteration 0 of the first
//                          operating point (op.go:25-88) on the OP-mode matrix; replayed here for the
//                          NOMINAL instance with MarkowitzLU (hostlu.hpp) and then frozen for the batch
//   elimination program    symbolic LU over the stamped pattern (OP + transient entries) in that order
#include <algorithm>
#include <cstring>
#include <set>
#include "tsb_internal.hpp"
#include "hostlu.hpp"
#include "device/models.cuh"

namespace tsb {

int device_num_outputs(const Dev& d) {
    switch (d.kind) {
    case TSB_R: return 1;
    case TSB_C: case TSB_L: case TSB_D: case TSB_LCORE: return 2;
    case TSB_Q: return 10;
    case TSB_M: return 22;
    case TSB_K: { int m = (int)d.ip.size(); return 3 * (m * (m - 1) / 2); }
    default: return 0;
    }
}

static int state_size(int kind) {
    switch (kind) {
    case TSB_C: return 2;      // V0, V1 (the charges q0 = C*V0, q1 = C*V1 are recomputed, models.cuh)
    case TSB_L: return 4;
    case TSB_D: return 1;
    case TSB_Q: return 3;
    case TSB_M: return 9;
    default: return 0;
    }
}

static const int OUT_CONST = -1, OUT_SRC = -2;

// The ordered AddElement / AddRHS calls of one device.  Entries whose row or column is ground are
// skipped exactly where the reference's `if n != 0` guards skip them.
void device_stamp_entries(const Plan& plan, int di, std::vector<StampEntry>& out) {
    const Dev& d = plan.devs[di];
    auto el = [&](int r, int c, int o, double sign, bool tran = false) {
        if (r != 0 && c != 0) out.push_back(StampEntry{r, c, o, sign, 0.0, tran});
    };
    auto elc = [&](int r, int c, double v) {
        if (r != 0 && c != 0) out.push_back(StampEntry{r, c, OUT_CONST, 1.0, v, false});
    };
    auto rhs = [&](int r, int o, double sign, bool tran = false) {
        if (r != 0) out.push_back(StampEntry{r, 0, o, sign, 0.0, tran});
    };
    const int* n = d.nodes;
    const int b = d.branch;
    switch (d.kind) {
    case TSB_R:       // resistor.go:60-71
        el(n[0], n[0], 0, +1); el(n[0], n[1], 0, -1); el(n[1], n[0], 0, -1); el(n[1], n[1], 0, +1);
        break;
    case TSB_C:       // capacitor.go:72-105
        el(n[0], n[0], 0, +1); el(n[0], n[1], 0, -1); rhs(n[0], 1, +1);
        el(n[1], n[1], 0, +1); el(n[1], n[0], 0, -1); rhs(n[1], 1, -1);
        break;
    case TSB_L: case TSB_LCORE:   // inductor.go:58-76, magnetic.go:206-250
        elc(n[0], b, -1.0); elc(b, n[0], -1.0); elc(n[1], b, 1.0); elc(b, n[1], 1.0);
        el(b, b, 0, +1); rhs(b, 1, +1);
        break;
    case TSB_V:       // vsource.go:139-151
        elc(b, n[0], 1.0); elc(n[0], b, 1.0); elc(b, n[1], -1.0); elc(n[1], b, -1.0);
        rhs(b, OUT_SRC, +1);
        break;
    case TSB_I:       // isource.go:138-145
        rhs(n[0], OUT_SRC, +1); rhs(n[1], OUT_SRC, -1);
        break;
    case TSB_D:       // diode.go:208-224
        el(n[0], n[0], 0, +1); el(n[0], n[1], 0, -1); rhs(n[0], 1, -1);
        el(n[1], n[0], 0, -1); el(n[1], n[1], 0, +1); rhs(n[1], 1, +1);
        break;
    case TSB_Q: {     // bjt.go:344-372 (collector, base, emitter)
        int nc = n[0], nb = n[1], ne = n[2];
        el(nc, nc, 0, +1); el(nc, nb, 1, +1); el(nc, ne, 2, +1); rhs(nc, 3, +1);
        el(nb, nb, 4, +1); el(nb, nc, 5, +1); rhs(nb, 6, +1);
        el(ne, ne, 7, +1); el(ne, nb, 8, +1); rhs(ne, 9, +1);
        break;
    }
    case TSB_M: {     // mosfet.go:701-783 (drain, gate, source, bulk)
        int nd = n[0], ng = n[1], ns = n[2], nb = n[3];
        el(nd, nd, 0, +1); el(nd, ng, 1, +1); el(nd, ns, 2, +1); el(nd, nb, 3, +1); rhs(nd, 4, +1);
        el(ns, ns, 5, +1); el(ns, nd, 6, +1); el(ns, ng, 7, +1); el(ns, nb, 8, +1); rhs(ns, 9, +1);
        if (ng != 0) {
            if (nd != 0) { el(ng, nd, 10, +1, true); el(nd, ng, 10, +1, true); rhs(ng, 11, +1, true); rhs(nd, 11, -1, true); }
            if (ns != 0) { el(ng, ns, 12, +1, true); el(ns, ng, 12, +1, true); rhs(ng, 13, +1, true); rhs(ns, 13, -1, true); }
            if (nb != 0) { el(ng, nb, 14, +1, true); el(nb, ng, 14, +1, true); rhs(ng, 15, +1, true); rhs(nb, 15, -1, true); }
            el(ng, ng, 16, +1, true);
        }
        if (nb != 0) {
            if (ns != 0) { el(nb, ns, 17, +1, true); el(ns, nb, 17, +1, true); rhs(nb, 18, +1, true); rhs(ns, 18, -1, true); }
            if (nd != 0) { el(nb, nd, 19, +1, true); el(nd, nb, 19, +1, true); rhs(nb, 20, +1, true); rhs(nd, 20, -1, true); }
            el(nb, nb, 21, +1, true);
        }
        break;
    }
    case TSB_K: {     // mutual.go:104-116
        int m = (int)d.ip.size(), q = 0;
        for (int i = 0; i < m; ++i)
            for (int j = i + 1; j < m; ++j, ++q) {
                int bi = plan.devs[d.ip[i]].branch, bj = plan.devs[d.ip[j]].branch;
                el(bi, bj, 3 * q, +1, true); el(bj, bi, 3 * q, +1, true);
                rhs(bi, 3 * q + 1, +1, true); rhs(bj, 3 * q + 2, +1, true);
            }
        break;
    }
    }
}

// StampAC of every device kind (AddComplexElement calls, ground rows / columns skipped like the reference's guards).
//   code 0: (+-g, 0)            resistor.go:41-54
//   code 1: (0, +-omega*C)      capacitor.go:48-66
//   code 2: (0, +-omega*L)      inductor.go:43-57 — an ADMITTANCE j*omega*L between the nodes; the branch row stays empty,
//                               so every circuit with an inductor is singular in the reference's AC analysis
//   code 3: (+-1, 0)            vsource.go:155-177 incidence
// Mutual and MagneticInductor stamp NOTHING in this mode: circuit.Stamp calls dev.Stamp (circuit.go:165-176), and their Stamp
// methods have no AC case (mutual.go:63-65 returns unless Mode == TransientAnalysis; magnetic.go:205-273 switches over OP and
// transient only) — the StampAC methods they define (mutual.go:122, magnetic.go:276) are never reached.  A core inductor's
// branch row is therefore empty as well.
// Nonlinear devices stamp small-signal values taken from the state their operating point left behind — which the
// reference computes on a COMPLEX matrix with a real-indexed right-hand side (matrix/circuit.go:99-105 vs :126-150):
// not reproducible without the un-vendored module's vector layout, so AC analysis is offered for linear circuits only.
void device_ac_entries(const Plan& plan, int di, std::vector<AcEntry>& out) {
    const Dev& d = plan.devs[di];
    auto quad = [&](int n1, int n2, int code) {          // (n1,n1) (n1,n2) | (n2,n1) (n2,n2) in the resistor's call order
        if (n1 != 0) { out.push_back({n1, n1, code, +1}); if (n2 != 0) out.push_back({n1, n2, code, -1}); }
        if (n2 != 0) { if (n1 != 0) out.push_back({n2, n1, code, -1}); out.push_back({n2, n2, code, +1}); }
    };
    auto quad_c = [&](int n1, int n2, int code) {        // capacitor / inductor order: (n1,n1) (n1,n2) | (n2,n2) (n2,n1)
        if (n1 != 0) { out.push_back({n1, n1, code, +1}); if (n2 != 0) out.push_back({n1, n2, code, -1}); }
        if (n2 != 0) { out.push_back({n2, n2, code, +1}); if (n1 != 0) out.push_back({n2, n1, code, -1}); }
    };
    const int* n = d.nodes;
    switch (d.kind) {
    case TSB_R: quad(n[0], n[1], 0); break;
    case TSB_C: quad_c(n[0], n[1], 1); break;
    case TSB_L: quad_c(n[0], n[1], 2); break;
    case TSB_V:
        if (n[0] != 0) { out.push_back({d.branch, n[0], 3, +1}); out.push_back({n[0], d.branch, 3, +1}); }
        if (n[1] != 0) { out.push_back({d.branch, n[1], 3, -1}); out.push_back({n[1], d.branch, 3, -1}); }
        break;
    default: break;
    }
}

// generateFrequencyPoints (ac.go:100-126) with Go's math.Log10 / Log2 / Pow: log2(x) = Log(frac)*(1/Ln2) + exp over Frexp's
// split, log10(x) = log2(x)*(Ln2/Ln10) (src/math/log10.go), Pow as restated in models.cuh.  A single point divides by zero
// there (step = NaN or Inf, the one frequency NaN); kept.
void ac_frequency_points(int sweep_type, int n_points, double fstart, double fstop, std::vector<double>& f) {
    auto go_log2 = [](double x) {
        int e;
        const double frac = std::frexp(x, &e);
        if (frac == 0.5) return (double)(e - 1);
        return std::log(frac) * (1 / 0.693147180559945309417232121458176568) + (double)e;
    };
    auto go_log10 = [&](double x) { return go_log2(x) * (0.693147180559945309417232121458176568 / 2.30258509299404568401799145468436421); };
    f.assign((size_t)n_points, 0.0);
    if (sweep_type == 0) {
        const double ls = go_log10(fstart), le = go_log10(fstop), step = (le - ls) / (double)(n_points - 1);
        for (int i = 0; i < n_points; ++i) f[i] = tsb_go_pow(10.0, ls + (double)i * step);
    } else if (sweep_type == 1) {
        const double ls = go_log2(fstart), le = go_log2(fstop), step = (le - ls) / (double)(n_points - 1);
        for (int i = 0; i < n_points; ++i) f[i] = tsb_go_pow(2.0, ls + (double)i * step);
    } else {
        const double step = (fstop - fstart) / (double)(n_points - 1);
        for (int i = 0; i < n_points; ++i) f[i] = fstart + (double)i * step;
    }
}

int Plan::num_columns(int an) const {
    if (an == TSB_AN_OP) return n_nodes + n_branches;
    if (an == TSB_AN_AC) {            // FREQ, then magnitude and phase of every node voltage and voltage-source current (ac.go:76-95)
        int nv = 0;
        for (const Dev& d : devs) if (d.kind == TSB_V) ++nv;
        return 1 + 2 * (n_nodes + nv);
    }
    int nr = 0;
    for (const Dev& d : devs) if (d.kind == TSB_R) ++nr;
    return (an == TSB_AN_DC2 ? 2 : 1) + n_nodes + n_branches + nr;     // nested sweep: SWEEP1, SWEEP2 (dc.go:272-288)
}

std::string Plan::column_name(int an, int col) const {
    int k = col;
    if (an == TSB_AN_AC) {            // StoreACResult (anlysis.go:87-111): <key>_MAG, <key>_PHASE
        if (k == 0) return "FREQ";
        --k;
        const char* suffix = (k & 1) ? "_PHASE" : "_MAG";
        k /= 2;
        if (k < n_nodes) return ((int)node_names.size() > k + 1 ? "V(" + node_names[k + 1] + ")" : "V(" + std::to_string(k + 1) + ")") + suffix;
        k -= n_nodes;
        for (const Dev& d : devs) if (d.kind == TSB_V) { if (k == 0) return "I(" + d.name + ")" + suffix; --k; }
        return "?";
    }
    if (an != TSB_AN_OP) {
        if (k == 0) return an == TSB_AN_TRAN ? "TIME" : "SWEEP1";
        --k;
        if (an == TSB_AN_DC2) {
            if (k == 0) return "SWEEP2";
            --k;
        }
    }
    if (k < n_nodes) {
        if ((int)node_names.size() > k + 1) return "V(" + node_names[k + 1] + ")";
        return "V(" + std::to_string(k + 1) + ")";
    }
    k -= n_nodes;
    if (k < n_branches) {
        for (const Dev& d : devs) if (d.branch == n_nodes + 1 + k) return "I(" + d.name + ")";
        return "I(b" + std::to_string(k + 1) + ")";
    }
    k -= n_branches;
    for (const Dev& d : devs) if (d.kind == TSB_R) { if (k == 0) return "I(" + d.name + ")"; --k; }
    return "?";
}

namespace {

// ---- host evaluation of the NOMINAL instance (symbolic pass only) ---------------------------
struct Nominal {
    const Plan& pl;
    std::vector<double> P, S, D, SV;
    explicit Nominal(const Plan& p) : pl(p), P(p.nominal), S(p.n_state, 0.0), D(std::max(1, p.n_derived), 0.0), SV(std::max(1, p.n_src), 0.0) {}

    void derive() {
        for (int di : pl.stamp_order) {
            const Dev& d = pl.devs[di];
            if (d.kind == TSB_R) D[d.d_off] = tsb_res_g(&P[d.p_off]);
            else if (d.kind == TSB_L) tsb_ind_derive(&P[d.p_off], &D[d.d_off]);
            else if (d.kind == TSB_LCORE) D[d.d_off] = tsb_lcore_L0(&P[d.p_off]);
            else if (d.kind == TSB_D) tsb_dio_derive(&P[d.p_off], &D[d.d_off]);
            else if (d.kind == TSB_M) tsb_mos_init_state(&P[d.p_off], &S[d.s_off]);
            else if (d.kind == TSB_K) {
                int m = (int)d.ip.size(), q = 0;
                for (int i = 0; i < m; ++i)
                    for (int j = i + 1; j < m; ++j, ++q)
                        D[d.d_off + q] = tsb_mut_M(P[d.p_off], ind_value(d.ip[i]), ind_value(d.ip[j]));
            }
        }
    }
    double ind_value(int di) const {
        const Dev& d = pl.devs[di];
        return d.kind == TSB_LCORE ? D[d.d_off] : P[d.p_off];
    }
    double ind_current(int di) const {
        const Dev& d = pl.devs[di];
        return d.kind == TSB_LCORE ? 0.0 : S[d.s_off];
    }
    void sources(double t, double fac) {
        for (const Dev& d : pl.devs) {
            if (d.src_slot < 0) continue;
            const double* p = &P[d.p_off];
            double v = 0;
            switch (d.src_type()) {
            case TSB_SRC_DC: v = p[0] * fac; break;
            case TSB_SRC_SIN: v = tsb_src_sin(p, t, fac); break;
            case TSB_SRC_PULSE: v = tsb_src_pulse(p, t); break;
            case TSB_SRC_PWL: v = tsb_src_pwl(p, (int)d.p.size() / 2, t); break;
            }
            SV[d.src_slot] = v;
        }
    }
    void eval(int di, const TsbEnv& e, double* o) {
        const Dev& d = pl.devs[di];
        const double* p = d.p.empty() ? nullptr : &P[d.p_off];
        double* s = d.n_state ? &S[d.s_off] : nullptr;
        switch (d.kind) {
        case TSB_R: o[0] = D[d.d_off]; break;
        case TSB_C: tsb_cap_eval(p, s, e, o); break;
        case TSB_L: tsb_ind_eval(p, s, e, o); break;
        case TSB_LCORE: tsb_lcore_eval(D[d.d_off], e, o); break;
        case TSB_D: tsb_dio_eval(p, &D[d.d_off], s, e, o); break;
        case TSB_Q: tsb_bjt_eval(p, s, d.ip.empty() ? 0 : d.ip[0], o); break;
        case TSB_M: tsb_mos_eval(p, s, d.ip.size() > 0 ? d.ip[0] : 1, d.ip.size() > 1 ? d.ip[1] : 0, e, o); break;
        case TSB_K: {
            int m = (int)d.ip.size(), q = 0;
            for (int i = 0; i < m; ++i)
                for (int j = i + 1; j < m; ++j, ++q)
                    tsb_mut_eval(D[d.d_off + q], ind_current(d.ip[i]), ind_current(d.ip[j]), e, o + 3 * q);
            break;
        }
        default: break;
        }
    }
    void update_nl(const std::vector<double>& x) {
        for (const Dev& d : pl.devs) {
            if (!d.nonlinear()) continue;
            double* s = &S[d.s_off];
            if (d.kind == TSB_D) s[0] = x[d.nodes[0]] - x[d.nodes[1]];
            else if (d.kind == TSB_Q) tsb_bjt_update(s, d.ip.empty() ? 0 : d.ip[0], x[d.nodes[0]], x[d.nodes[1]], x[d.nodes[2]]);
            else tsb_mos_update(s, d.ip.size() > 1 ? d.ip[1] : 0, x[d.nodes[0]], x[d.nodes[1]], x[d.nodes[2]], x[d.nodes[3]]);
        }
    }
    // Stamp into a MarkowitzLU + rhs; `linear_only` = calculateInitialEstimate (op.go:90-111).
    void stamp(MarkowitzLU& m, std::vector<double>& b, const TsbEnv& e, bool linear_only, bool op_calls_only) {
        double o[64];
        for (int di : pl.stamp_order) {
            const Dev& d = pl.devs[di];
            if (linear_only && d.nonlinear()) continue;
            std::vector<double> big;
            double* op = o;
            int no = device_num_outputs(d);
            if (no > 64) { big.resize(no); op = big.data(); }
            eval(di, e, op);
            for (const StampEntry& s : pl.stamps[di]) {
                if (op_calls_only && s.tran_only) continue;
                double v = s.out == OUT_CONST ? s.cval : (s.out == OUT_SRC ? SV[d.src_slot] : op[s.out]);
                if (s.col == 0) {
                    if (d.kind == TSB_C || d.kind == TSB_LCORE) { if (e.mode != TSB_MODE_TRAN) continue; }   // no AddRHS in OP mode
                    b[s.row] += s.sign * v;
                } else m.add(s.row, s.col, s.sign * v);
            }
        }
    }
};

void build_lu_program(int n, const std::vector<std::pair<int, int>>& pattern, const PivotOrder& ord, bool dense, LuProgram& lu) {
    lu = LuProgram();
    lu.n = n; lu.prow = ord.prow; lu.pcol = ord.pcol; lu.dense = dense;
    auto add = [&](int r, int c) {
        auto key = std::make_pair(r, c);
        auto it = lu.index.find(key);
        if (it != lu.index.end()) return it->second;
        int k = (int)lu.pos.size();
        lu.pos.push_back(key); lu.index[key] = k;
        return k;
    };
    if (dense) { for (int r = 1; r <= n; ++r) for (int c = 1; c <= n; ++c) add(r, c); }
    else for (auto& rc : pattern) add(rc.first, rc.second);
    std::vector<int> rstep(n + 1, 0), cstep(n + 1, 0);
    for (int k = 1; k <= n; ++k) { rstep[ord.prow[k]] = k; cstep[ord.pcol[k]] = k; }
    lu.steps.assign(n + 1, LuProgram::Step());
    for (int k = 1; k <= n; ++k) {
        LuProgram::Step& st = lu.steps[k];
        int pr = ord.prow[k], pc = ord.pcol[k];
        st.piv = add(pr, pc);
        std::vector<std::pair<int, int>> us, ls;      // (step of col/row, ext index)
        for (int c = 1; c <= n; ++c) if (cstep[c] > k && lu.index.count({pr, c})) us.push_back({cstep[c], c});
        for (int r = 1; r <= n; ++r) if (rstep[r] > k && lu.index.count({r, pc})) ls.push_back({rstep[r], r});
        std::sort(us.begin(), us.end()); std::sort(ls.begin(), ls.end());
        for (auto& u : us) { st.urow.push_back(lu.index[{pr, u.second}]); st.ucol_step.push_back(u.first); }
        for (auto& l : ls) { st.lcol.push_back(lu.index[{l.second, pc}]); st.lrow_step.push_back(l.first); }
        st.target.assign(us.size(), std::vector<int>(ls.size(), -1));
        for (size_t ui = 0; ui < us.size(); ++ui)
            for (size_t li = 0; li < ls.size(); ++li)
                st.target[ui][li] = add(ls[li].second, us[ui].second);     // creates fill-in
    }
}

// ---- static condensation order for the transient solves of the fast build ------------------------------------------
// In a transient run most of the MNA matrix is the same in every solve: resistor conductances and the +-1 incidence
// entries of sources and inductor branches do not depend on the time step or on device state; only the companion
// conductances C/dt, L/dt, M/dt (and every entry a nonlinear device stamps) change.  The reference's order (frozen from
// its operating-point matrix) happens to start with an inductor branch diagonal in the bundled decks, which makes every
// later operation depend on dt.  Here the pivots whose value is invariant are eliminated FIRST: all their
// multipliers, and every update between invariant entries, are computed once per instance (Ckt::prefactor), and a solve
// only factors the Schur complement of the variant entries.  Candidates are taken in Markowitz order (fewest
// operations; the diagonal on ties), numerically acceptable against the invariant entries of their column on the nominal
// instance (relative threshold 1e-3 as in Sparse 1.3); the variant remainder is ordered by MarkowitzLU on the nominal
// transient matrix at a representative step.  A different elimination order is a different rounding pattern — which is
// what the fast build is allowed (its results are held to the 1e-9 / 1e-12 contract by the parity tests on every deck
// but the ill-conditioned coupled-inductor ones); the strict build keeps the reference's order.
static bool stamp_is_variant(const Dev& d, const StampEntry& s) {
    if (s.col == 0) return true;                                   // right-hand side: sources and state
    switch (d.kind) {
    case TSB_R: case TSB_V: case TSB_I: return false;
    case TSB_L: case TSB_LCORE: return s.out != OUT_CONST;        // the branch diagonal -L/dt
    default: return true;                                          // C, K, D, Q, M
    }
}

static void build_tranfast(Plan& pl, Nominal& nom) {
    const int n = pl.n();
    pl.has_tranfast = false;
    if (pl.has_bjt || n > 48) return;                               // BJT circuits stay dense in the reference's order (Inf / NaN bookkeeping)
    const double dt_rep = 1e-6;
    std::vector<std::vector<char>> on(n + 1, std::vector<char>(n + 1, 0)), var(n + 1, std::vector<char>(n + 1, 0));
    std::vector<std::vector<double>> val(n + 1, std::vector<double>(n + 1, 0.0));
    {
        TsbEnv env{TSB_MODE_TRAN, 0.0, dt_rep, 0.0, 1.0 / dt_rep};
        double o[64];
        for (int di : pl.stamp_order) {
            const Dev& d = pl.devs[di];
            std::vector<double> big;
            double* op = o;
            int no = device_num_outputs(d);
            if (no > 64) { big.resize(no); op = big.data(); }
            nom.eval(di, env, op);
            for (const StampEntry& s : pl.stamps[di]) {
                if (s.col == 0) continue;
                on[s.row][s.col] = 1;
                if (stamp_is_variant(d, s)) var[s.row][s.col] = 1;
                double v = s.out == OUT_CONST ? s.cval : op[s.out];
                val[s.row][s.col] += s.sign * v;
            }
        }
    }
    // entry flags by stamp, before any elimination: what the code generator needs
    std::vector<std::vector<char>> var0 = var;
    std::vector<char> rdone(n + 1, 0), cdone(n + 1, 0);
    PivotOrder ord; ord.n = n; ord.ext2int.assign(n + 1, 0); ord.prow.assign(n + 1, 0); ord.pcol.assign(n + 1, 0);
    int k = 0;
    for (;;) {
        int br = 0, bc = 0; long best = -1; double bmag = 0;
        for (int r = 1; r <= n; ++r) {
            if (rdone[r]) continue;
            for (int c = 1; c <= n; ++c) {
                if (cdone[c] || !on[r][c] || var[r][c] || val[r][c] == 0.0) continue;
                double colmax = 0;
                int rc = 0, cc = 0;
                for (int i = 1; i <= n; ++i) if (!rdone[i] && on[i][c]) { ++cc; if (!var[i][c]) colmax = std::max(colmax, std::fabs(val[i][c])); }
                for (int j = 1; j <= n; ++j) if (!cdone[j] && on[r][j]) ++rc;
                if (std::fabs(val[r][c]) < 1e-3 * colmax) continue;
                // a variant entry in the pivot's column or row may be arbitrarily large: only rows / columns whose other
                // live entries are all invariant qualify, except that the +-1 incidence pivots of sources are exact
                // (a row or column singleton needs no such care: eliminating it updates no other matrix entry at all)
                long cost = (long)(rc - 1) * (cc - 1);
                bool clean = true;
                for (int i = 1; i <= n; ++i) if (i != r && !rdone[i] && on[i][c] && var[i][c]) clean = false;
                for (int j = 1; j <= n; ++j) if (j != c && !cdone[j] && on[r][j] && var[r][j]) clean = false;
                if (!clean && cost != 0) continue;
                double mag = std::fabs(val[r][c]);
                bool better = best < 0 || cost < best || (cost == best && ((r == c) > (br == bc))) ||
                              (cost == best && (r == c) == (br == bc) && mag > bmag);
                if (better) { best = cost; br = r; bc = c; bmag = mag; }
            }
        }
        if (!br) break;
        ++k; ord.prow[k] = br; ord.pcol[k] = bc; rdone[br] = 1; cdone[bc] = 1;
        for (int i = 1; i <= n; ++i) {
            if (rdone[i] || !on[i][bc]) continue;
            for (int j = 1; j <= n; ++j) {
                if (cdone[j] || !on[br][j]) continue;
                on[i][j] = 1;
                var[i][j] = var[i][j] | var[i][bc] | var[br][j];
                val[i][j] -= val[i][bc] * val[br][j] / val[br][bc];
            }
        }
    }
    const int n_inv = k;
    if (n_inv == 0) return;                                         // nothing to condense
    if (k < n) {                                                    // the variant remainder: Markowitz / threshold on the nominal values
        std::vector<int> rows, cols;
        for (int r = 1; r <= n; ++r) if (!rdone[r]) rows.push_back(r);
        for (int c = 1; c <= n; ++c) if (!cdone[c]) cols.push_back(c);
        const int m = (int)rows.size();
        MarkowitzLU rem(m);
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) if (on[rows[i]][cols[j]]) rem.add(i + 1, j + 1, val[rows[i]][cols[j]]);
        if (rem.assigned() != m || !rem.order_and_factor()) return;
        for (int q = 1; q <= m; ++q) { ++k; ord.prow[k] = rows[rem.pivot_row(q) - 1]; ord.pcol[k] = cols[rem.pivot_col(q) - 1]; }
    }
    std::vector<std::pair<int, int>> pat = pl.pattern_op;
    pat.insert(pat.end(), pl.pattern_tran_extra.begin(), pl.pattern_tran_extra.end());
    build_lu_program(n, pat, ord, false, pl.lu_tf);
    pl.tf_variant.assign(pl.lu_tf.pos.size(), 0);
    for (size_t q = 0; q < pl.lu_tf.pos.size(); ++q) pl.tf_variant[q] = var0[pl.lu_tf.pos[q].first][pl.lu_tf.pos[q].second];
    pl.has_tranfast = true;
}


// ---- cooperative mapping: partition + nested-dissection order (CoopPlan, tsb_internal.hpp) ---------------------------------
// Graph: one vertex per unknown, a clique per device over the unknowns it stamps or reads.  Recursive bisection by
// breadth-first level sets (every root, every level tried; the cheapest separator that leaves two non-empty sides wins,
// ties to the better balance), then a pass that returns separator vertices with neighbours on one side only to that side.
// The graphs are tiny (n <= 64), so exhaustive search over roots and levels costs nothing.
namespace coop_detail {

struct Cut { std::vector<int> A, B, S; bool ok = false; };

Cut bisect(const std::vector<int>& V, const std::vector<std::vector<char>>& adj, int n) {
    Cut best; long best_score = -1;
    std::vector<char> in(n + 1, 0);
    for (int v : V) in[v] = 1;
    for (int root : V) {
        std::vector<int> level(n + 1, -1), q{root};
        level[root] = 0;
        int k = 0;
        for (size_t h = 0; h < q.size(); ++h) {
            int v = q[h];
            for (int w = 1; w <= n; ++w) if (in[w] && adj[v][w] && level[w] < 0) { level[w] = level[v] + 1; k = std::max(k, level[w]); q.push_back(w); }
        }
        for (int i = 0; i <= k; ++i) {          // i = 0: only meaningful when the subgraph is disconnected (empty separator)
            std::vector<int> side(n + 1, 0);     // 1 A, 2 B, 3 S
            int nA = 0, nB = 0;
            for (int v : V) {
                if (level[v] < 0) continue;
                if (i == 0) { side[v] = 1; ++nA; }
                else if (level[v] < i) { side[v] = 1; ++nA; }
                else if (level[v] == i) side[v] = 3;
                else { side[v] = 2; ++nB; }
            }
            for (int v : V) if (level[v] < 0) { if (nA <= nB) { side[v] = 1; ++nA; } else { side[v] = 2; ++nB; } }   // other components
            for (int v : V) {                     // a separator vertex that touches one side only belongs to that side
                if (side[v] != 3) continue;
                bool tA = false, tB = false;
                for (int w : V) if (adj[v][w]) { if (side[w] == 1) tA = true; if (side[w] == 2) tB = true; }
                if (tA && tB) continue;
                if (tA || (!tB && nA <= nB)) { side[v] = 1; ++nA; } else { side[v] = 2; ++nB; }
            }
            if (nA == 0 || nB == 0) continue;
            int nS = 0;
            for (int v : V) if (side[v] == 3) ++nS;
            bool clean = true;                    // (the moves above cannot connect A and B, checked anyway)
            for (int v : V) for (int w : V) if (adj[v][w] && side[v] == 1 && side[w] == 2) clean = false;
            if (!clean) continue;
            long score = 2L * nS + std::labs((long)nA - nB);
            if (best_score < 0 || score < best_score) {
                best_score = score; best = Cut(); best.ok = true;
                for (int v : V) (side[v] == 1 ? best.A : side[v] == 2 ? best.B : best.S).push_back(v);
            }
        }
    }
    return best;
}

}  // namespace coop_detail

static bool build_coop(const Plan& pl, Nominal& nom, int parts, CoopPlan& cp) {
    const int n = pl.n();
    if (pl.has_bjt || pl.has_mutual || n < 2 * parts || n > 96 || (parts != 2 && parts != 4 && parts != 8)) return false;     // BJT circuits are generated dense (NaN bookkeeping): no partition
    // ---- graph ------------------------------------------------------------------------------------------------------------
    std::vector<std::vector<char>> adj(n + 1, std::vector<char>(n + 1, 0));
    std::vector<std::vector<int>> dev_unk(pl.devs.size());
    for (size_t di = 0; di < pl.devs.size(); ++di) {
        const Dev& d = pl.devs[di];
        std::set<int> u;
        for (const StampEntry& s : pl.stamps[di]) { if (s.row) u.insert(s.row); if (s.col) u.insert(s.col); }
        for (int i = 0; i < d.n_nodes; ++i) if (d.nodes[i]) u.insert(d.nodes[i]);
        if (d.branch > pl.n_nodes) u.insert(d.branch);
        dev_unk[di].assign(u.begin(), u.end());
        for (int a : u) for (int b : u) if (a != b) adj[a][b] = 1;
    }
    // ---- partition ---------------------------------------------------------------------------------------------------------
    cp = CoopPlan();
    cp.parts = parts;
    cp.owner.assign(n + 1, -1);
    std::vector<int> all;
    for (int v = 1; v <= n; ++v) all.push_back(v);
    std::vector<std::vector<int>> interior;
    {
        int levels = 0;
        while ((1 << levels) < parts) ++levels;
        // recursive bisection: the separators of all levels together are THE separator, the leaves the interiors
        struct Rec {
            const std::vector<std::vector<char>>& adj; int n; std::vector<std::vector<int>>& out;
            bool run(const std::vector<int>& V, int lv) {
                if (lv == 0) { out.push_back(V); return !V.empty(); }
                coop_detail::Cut c = coop_detail::bisect(V, adj, n);
                return c.ok && run(c.A, lv - 1) && run(c.B, lv - 1);
            }
        } rec{adj, n, interior};
        // hubs (a supply rail every stage hangs on) make every breadth-first level set huge: they go to the separator first
        std::vector<int> rest;
        for (int v : all) {
            int deg = 0;
            for (int w = 1; w <= n; ++w) deg += adj[v][w];
            if (deg >= std::max(6, n / 2)) continue;
            rest.push_back(v);
        }
        if (!rec.run(rest, levels) || (int)interior.size() != parts) return false;
    }
    for (int p = 0; p < parts; ++p) {
        if (interior[p].empty()) return false;
        std::sort(interior[p].begin(), interior[p].end());
        for (int v : interior[p]) cp.owner[v] = p;
    }
    // ---- the nominal transient matrix (pattern and values) ------------------------------------------------------------------
    const double dt_rep = 1e-6;
    std::vector<std::vector<char>> on(n + 1, std::vector<char>(n + 1, 0));
    std::vector<std::vector<double>> val(n + 1, std::vector<double>(n + 1, 0.0));
    {
        TsbEnv env{TSB_MODE_TRAN, 0.0, dt_rep, 0.0, 1.0 / dt_rep};
        double o[64];
        for (int di : pl.stamp_order) {
            const Dev& d = pl.devs[di];
            if (device_num_outputs(d) > 64) return false;
            nom.eval(di, env, o);
            for (const StampEntry& s : pl.stamps[di]) {
                if (s.col == 0) continue;
                on[s.row][s.col] = 1;
                val[s.row][s.col] += s.sign * (s.out == OUT_CONST ? s.cval : o[s.out]);
            }
        }
    }
    // an interior unknown whose row or column has no entry inside its own interior (a voltage source's branch whose node went
    // to the separator: its row holds the +-1 at that node only) cannot be pivoted there: it joins the separator
    for (bool changed = true; changed;) {
        changed = false;
        for (int u = 1; u <= n; ++u) {
            const int p = cp.owner[u];
            if (p < 0) continue;
            bool row = false, col = false;
            for (int v = 1; v <= n; ++v) if (cp.owner[v] == p) { row = row || on[u][v]; col = col || on[v][u]; }
            if (!row || !col) { cp.owner[u] = -1; changed = true; }
        }
    }
    for (int p = 0; p < parts; ++p) {
        interior[p].clear();
        for (int u = 1; u <= n; ++u) if (cp.owner[u] == p) interior[p].push_back(u);
        if (interior[p].empty()) return false;
    }
    // ---- devices and result columns ---------------------------------------------------------------------------------------
    cp.dev_owner.assign(pl.devs.size(), -1);
    std::vector<int> load(parts, 0);
    for (size_t di = 0; di < pl.devs.size(); ++di) {
        int own = -1;
        for (int u : dev_unk[di]) if (cp.owner[u] >= 0) { if (own >= 0 && own != cp.owner[u]) return false; own = cp.owner[u]; }
        cp.dev_owner[di] = own;
        if (own >= 0) ++load[own];
    }
    for (size_t di = 0; di < pl.devs.size(); ++di)           // devices between separator unknowns only: to the least loaded part
        if (cp.dev_owner[di] < 0) { int p = (int)(std::min_element(load.begin(), load.end()) - load.begin()); cp.dev_owner[di] = p; ++load[p]; }
    const int ncol = pl.num_columns(TSB_AN_TRAN);
    cp.col_owner.assign(ncol, 0);
    std::vector<int> ncols(parts, 0);
    ncols[0] = 1;                                              // TIME
    for (int u = 1; u <= n; ++u) if (cp.owner[u] >= 0) { cp.col_owner[u] = cp.owner[u]; ++ncols[cp.owner[u]]; }
    {
        int c = n + 1;
        for (size_t di = 0; di < pl.devs.size(); ++di) if (pl.devs[di].kind == TSB_R) { if (c >= ncol) return false; cp.col_owner[c] = cp.dev_owner[di]; ++ncols[cp.dev_owner[di]]; ++c; }
        if (c != ncol) return false;
    }
    for (int u = 1; u <= n; ++u) if (cp.owner[u] < 0) { int p = (int)(std::min_element(ncols.begin(), ncols.end()) - ncols.begin()); cp.col_owner[u] = p; ++ncols[p]; }
    // ---- nested-dissection order ------------------------------------------------------------------------------------------
    PivotOrder ord; ord.n = n; ord.ext2int.assign(n + 1, 0); ord.prow.assign(n + 1, 0); ord.pcol.assign(n + 1, 0);
    cp.step_owner.assign(n + 1, -1);
    std::vector<char> rdone(n + 1, 0), cdone(n + 1, 0);
    int k = 0;
    auto order_block = [&](const std::vector<int>& rows, const std::vector<int>& cols, int owner) -> bool {
        const int m = (int)rows.size();
        if (m == 0) return true;
        MarkowitzLU blk(m);
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) if (on[rows[i]][cols[j]]) blk.add(i + 1, j + 1, val[rows[i]][cols[j]]);
        if (blk.assigned() != m || !blk.order_and_factor()) return false;
        for (int q = 1; q <= m; ++q) {
            const int br = rows[blk.pivot_row(q) - 1], bc = cols[blk.pivot_col(q) - 1];
            ++k; ord.prow[k] = br; ord.pcol[k] = bc; cp.step_owner[k] = owner; rdone[br] = 1; cdone[bc] = 1;
            if (val[br][bc] == 0.0) return false;
            for (int i = 1; i <= n; ++i) {             // numeric elimination on the full matrix: the separator block becomes the Schur complement
                if (rdone[i] || !on[i][bc]) continue;
                for (int j = 1; j <= n; ++j) {
                    if (cdone[j] || !on[br][j]) continue;
                    on[i][j] = 1;
                    val[i][j] -= val[i][bc] * val[br][j] / val[br][bc];
                }
            }
        }
        return true;
    };
    for (int p = 0; p < parts; ++p) if (!order_block(interior[p], interior[p], p)) return false;
    cp.n_int = k;
    std::vector<int> sep;
    for (int v = 1; v <= n; ++v) if (cp.owner[v] < 0) sep.push_back(v);
    if (!order_block(sep, sep, -1)) return false;
    if (k != n) return false;
    std::vector<std::pair<int, int>> pat = pl.pattern_op;
    pat.insert(pat.end(), pl.pattern_tran_extra.begin(), pl.pattern_tran_extra.end());
    build_lu_program(n, pat, ord, false, cp.lu);
    for (const auto& rc : cp.lu.pos) {                       // no entry may couple two interiors (fill included)
        const int a = cp.owner[rc.first], b = cp.owner[rc.second];
        if (a >= 0 && b >= 0 && a != b) return false;
    }
    return true;
}

}  // namespace

int plan_finalize(Plan& pl) {
    pl.error.clear();
    const int n = pl.n();
    if (n <= 0) { pl.error = "empty circuit"; return TSB_E_INVALID; }
    // validate + stamp order
    pl.stamp_order.clear();
    for (size_t i = 0; i < pl.devs.size(); ++i) if (pl.devs[i].kind != TSB_K) pl.stamp_order.push_back((int)i);
    for (size_t i = 0; i < pl.devs.size(); ++i) if (pl.devs[i].kind == TSB_K) pl.stamp_order.push_back((int)i);
    static const int need_nodes[10] = {2, 2, 2, 2, 2, 2, 3, 4, 0, 2};
    static const int need_p[10] = {1, 1, 1, 1, 1, 3, 9, 29, 1, 3};
    pl.n_params = pl.n_state = pl.n_src = pl.n_derived = 0;
    pl.has_nonlinear = pl.has_time_dependent = pl.has_bjt = pl.has_mutual = false;
    pl.nominal.clear();
    for (Dev& d : pl.devs) {
        if (d.kind < 0 || d.kind > 9) { pl.error = "bad device kind on " + d.name; return TSB_E_INVALID; }
        // the reference panics on a wrong node count (diode.go:45-47, bjt.go:69-71, mosfet.go:126-128)
        if (d.n_nodes != need_nodes[d.kind]) { pl.error = "device " + d.name + ": wrong number of nodes"; return TSB_E_INVALID; }
        if ((int)d.p.size() < need_p[d.kind]) { pl.error = "device " + d.name + ": too few parameters"; return TSB_E_INVALID; }
        for (int i = 0; i < d.n_nodes; ++i)
            if (d.nodes[i] < 0 || d.nodes[i] > pl.n_nodes) { pl.error = "device " + d.name + ": node index out of range"; return TSB_E_INVALID; }
        bool needs_branch = d.kind == TSB_L || d.kind == TSB_V || d.kind == TSB_LCORE;
        if (needs_branch && (d.branch <= pl.n_nodes || d.branch > n)) { pl.error = "device " + d.name + ": branch index out of range"; return TSB_E_INVALID; }
        if (d.kind == TSB_V || d.kind == TSB_I) {
            int st = d.src_type();
            size_t need = st == TSB_SRC_DC ? 1 : st == TSB_SRC_SIN ? 4 : st == TSB_SRC_PULSE ? 7 : 4;
            if (st < 0 || st > 3 || d.p.size() < need || (st == TSB_SRC_PWL && d.p.size() % 2)) { pl.error = "device " + d.name + ": bad source waveform"; return TSB_E_INVALID; }
            d.src_slot = pl.n_src++;
        } else d.src_slot = -1;
        if (d.kind == TSB_K) {
            if (d.ip.size() < 2) { pl.error = "mutual coupling " + d.name + " requires at least two inductors"; return TSB_E_INVALID; }
            for (int idx : d.ip)
                if (idx < 0 || idx >= (int)pl.devs.size() || (pl.devs[idx].kind != TSB_L && pl.devs[idx].kind != TSB_LCORE)) {
                    pl.error = "mutual coupling " + d.name + ": not an inductor"; return TSB_E_INVALID;
                }
        }
        d.p_off = pl.n_params; pl.n_params += (int)d.p.size();
        pl.nominal.insert(pl.nominal.end(), d.p.begin(), d.p.end());
        d.n_state = state_size(d.kind); d.s_off = pl.n_state; pl.n_state += d.n_state;
        d.d_off = -1;
        if (d.kind == TSB_R || d.kind == TSB_LCORE) { d.d_off = pl.n_derived; pl.n_derived += 1; }
        if (d.kind == TSB_L || d.kind == TSB_D) { d.d_off = pl.n_derived; pl.n_derived += 3; }
        if (d.kind == TSB_K) { int m = (int)d.ip.size(); d.d_off = pl.n_derived; pl.n_derived += m * (m - 1) / 2; }
        if (d.nonlinear()) pl.has_nonlinear = true;
        if (d.time_dependent()) pl.has_time_dependent = true;
        if (d.kind == TSB_Q) pl.has_bjt = true;
        if (d.kind == TSB_K) pl.has_mutual = true;
    }
    pl.stamps.assign(pl.devs.size(), {});
    for (size_t i = 0; i < pl.devs.size(); ++i) device_stamp_entries(pl, (int)i, pl.stamps[i]);

    // stamped patterns in first-touch order
    pl.pattern_op.clear(); pl.pattern_tran_extra.clear();
    {
        std::set<std::pair<int, int>> seen;
        for (int di : pl.stamp_order)
            for (const StampEntry& s : pl.stamps[di])
                if (s.col != 0 && !s.tran_only && seen.insert({s.row, s.col}).second) pl.pattern_op.push_back({s.row, s.col});
        for (int di : pl.stamp_order)
            for (const StampEntry& s : pl.stamps[di])
                if (s.col != 0 && s.tran_only && seen.insert({s.row, s.col}).second) pl.pattern_tran_extra.push_back({s.row, s.col});
    }

    // ---- replay the reference's first operating point for the nominal instance ----------------
    Nominal nom(pl);
    nom.derive();
    TsbEnv env{TSB_MODE_OP, 0.0, 0.0, 0.0, 0.0};
    {   // SetupDevices' initial stamp (circuit.go:154-156): only its side effects on device state matter
        double o[64]; std::vector<double> big;
        for (int di : pl.stamp_order) {
            const Dev& d = pl.devs[di];
            if (!d.nonlinear()) continue;
            int no = device_num_outputs(d);
            double* op = o; if (no > 64) { big.resize(no); op = big.data(); }
            nom.eval(di, env, op);
        }
    }
    nom.sources(0.0, 1.0);
    // calculateInitialEstimate: fresh sparse matrix, linear devices only
    std::vector<double> x0(n + 1, 0.0);
    {
        MarkowitzLU mi(n);
        std::vector<double> b(n + 1, 0.0);
        nom.stamp(mi, b, env, true, true);
        std::vector<std::pair<int, int>> pat;
        {
            std::set<std::pair<int, int>> seen;
            for (int di : pl.stamp_order) {
                if (pl.devs[di].nonlinear()) continue;
                for (const StampEntry& s : pl.stamps[di])
                    if (s.col != 0 && !s.tran_only && seen.insert({s.row, s.col}).second) pat.push_back({s.row, s.col});
            }
        }
        pl.order_init = PivotOrder();
        pl.order_init.n = n;
        pl.order_init.ext2int.assign(n + 1, 0);
        pl.order_init.prow.assign(n + 1, 0); pl.order_init.pcol.assign(n + 1, 0);
        bool ok = mi.assigned() == n && mi.order_and_factor();
        pl.order_init.singular = !ok;
        pl.init_struct_singular = !ok;
        if (ok) {
            for (int i = 1; i <= n; ++i) { pl.order_init.ext2int[i] = mi.ext2int(i); pl.order_init.prow[i] = mi.pivot_row(i); pl.order_init.pcol[i] = mi.pivot_col(i); }
            mi.solve(b, x0);
            nom.update_nl(x0);
            build_lu_program(n, pat, pl.order_init, false, pl.lu_init);
        }
    }
    // iteration 0 of OperatingPoint.doNRiter: UpdateNonlinearVoltages(x0), Stamp, Solve -> first Factor()
    {
        nom.update_nl(x0);
        MarkowitzLU mm(n);
        // Translate numbering = setup-stamp call order, then SetupElements touches every (i, j)
        for (int di : pl.stamp_order)
            for (const StampEntry& s : pl.stamps[di])
                if (s.col != 0 && !s.tran_only) mm.create(s.row, s.col);
        for (int i = 1; i <= n; ++i) for (int j = 1; j <= n; ++j) mm.create(i, j);
        std::vector<double> b(n + 1, 0.0);
        nom.stamp(mm, b, env, false, true);
        pl.order_main = PivotOrder();
        pl.order_main.n = n;
        pl.order_main.ext2int.assign(n + 1, 0);
        pl.order_main.prow.assign(n + 1, 0); pl.order_main.pcol.assign(n + 1, 0);
        for (int i = 1; i <= n; ++i) pl.order_main.ext2int[i] = mm.ext2int(i);
        if (!mm.order_and_factor()) {
            pl.order_main.singular = true;
            pl.error = "matrix is singular at the nominal operating point (no pivot order)";
            return TSB_E_INVALID;
        }
        for (int i = 1; i <= n; ++i) { pl.order_main.prow[i] = mm.pivot_row(i); pl.order_main.pcol[i] = mm.pivot_col(i); }
    }
    {
        std::vector<std::pair<int, int>> pat = pl.pattern_op;
        pat.insert(pat.end(), pl.pattern_tran_extra.begin(), pl.pattern_tran_extra.end());
        // BJT circuits run dense so that Inf/NaN propagate through the solve exactly as they do
        // through the reference's structurally dense matrix.
        build_lu_program(n, pat, pl.order_main, pl.has_bjt, pl.lu_main);
    }
    {   // AC analysis: same order, pattern = what StampAC touches
        std::vector<std::pair<int, int>> pat;
        std::set<std::pair<int, int>> seen;
        for (int di : pl.stamp_order) {
            std::vector<AcEntry> ae;
            device_ac_entries(pl, di, ae);
            for (const AcEntry& a : ae) if (seen.insert({a.row, a.col}).second) pat.push_back({a.row, a.col});
        }
        // the pivots of the frozen order exist in the reference's (structurally dense) matrix whether StampAC touches them or
        // not: an untouched pivot is an exact zero, "matrix factorization failed"
        for (int k = 1; k <= n; ++k) if (seen.insert({pl.order_main.prow[k], pl.order_main.pcol[k]}).second) pat.push_back({pl.order_main.prow[k], pl.order_main.pcol[k]});
        build_lu_program(n, pat, pl.order_main, false, pl.lu_ac);
    }
    {
        Nominal nom2(pl);            // fresh device state: the order must not depend on where the replay above left it
        nom2.derive();
        build_tranfast(pl, nom2);
    }
    pl.coop.clear();
    for (int parts : {2, 4, 8}) {
        Nominal nom3(pl);
        nom3.derive();
        CoopPlan cp;
        if (build_coop(pl, nom3, parts, cp)) { coop_dimensions(pl, cp, cp.nx, cp.nown_max); pl.coop[parts] = cp; }
    }
    pl.finalized = true;
    return TSB_OK;
}

}  // namespace tsb
