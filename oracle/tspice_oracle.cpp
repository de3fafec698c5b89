// ORACLE — TEST INFRASTRUCTURE ONLY (see sparse13.hpp header).  PARITY UNPINNED.
//
// tspice_oracle.cpp — flat C API over engine.hpp so that tests/ and bench.py's cpu_baseline /
// --impl reference legs can drive the CPU restatement through ctypes: build a circuit from a
// device table (the same table the product's C-ABI takes), then run OP / DC sweep / transient
// for N independent instances with per-instance parameter overrides, one task per instance on
// a std::thread pool ("one goroutine per instance").
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include "engine.hpp"

using namespace orc;

namespace {

enum Kind { K_R = 0, K_C = 1, K_L = 2, K_V = 3, K_I = 4, K_D = 5, K_Q = 6, K_M = 7, K_K = 8, K_LCORE = 9 };

struct DevDesc {
    int kind = 0;
    std::string name;
    int nodes[4] = {0, 0, 0, 0};
    int nn = 0, branch = 0;
    std::vector<double> p;
    std::vector<int> ip;
};

struct Template {
    int n_nodes = 0, n_branches = 0;
    std::vector<DevDesc> devs;
};

void fill_waveform(Waveform& w, const DevDesc& d) {
    w.stype = d.ip.empty() ? SRC_DC : d.ip[0];
    const std::vector<double>& p = d.p;
    switch (w.stype) {
    case SRC_DC: w.dcValue = p[0]; break;
    case SRC_SIN: w.dcValue = p[0]; w.amplitude = p[1]; w.freq = p[2]; w.phase = p[3]; break;
    case SRC_PULSE: w.v1 = p[0]; w.v2 = p[1]; w.delay = p[2]; w.rise = p[3]; w.fall = p[4]; w.pWidth = p[5]; w.period = p[6]; break;
    case SRC_PWL:
        for (size_t i = 0; i + 1 < p.size(); i += 2) { w.times.push_back(p[i]); w.values.push_back(p[i + 1]); }
        break;
    }
}

// netlist.CreateDevice + circuit.SetupDevices (parser.go:752-915, circuit.go:78-163)
std::unique_ptr<Circuit> build(const Template& t, const std::vector<DevDesc>& devs) {
    std::unique_ptr<Circuit> c(new Circuit);
    c->numNodes = t.n_nodes; c->numBranches = t.n_branches;
    c->CreateMatrix();
    std::vector<Device*> by_index(devs.size(), nullptr);
    for (int pass = 0; pass < 2; ++pass) {
        for (size_t k = 0; k < devs.size(); ++k) {
            const DevDesc& d = devs[k];
            if ((d.kind == K_K) != (pass == 1)) continue;
            std::unique_ptr<Device> dev;
            switch (d.kind) {
            case K_R: { auto* r = new Resistor; r->Value = d.p[0]; dev.reset(r); break; }
            case K_C: { auto* x = new Capacitor; x->Value = d.p[0]; dev.reset(x); break; }
            case K_L: { auto* x = new Inductor; x->Value = d.p[0]; x->branchIdx = d.branch; dev.reset(x); break; }
            case K_V: {
                auto* x = new VSource; fill_waveform(x->w, d); x->branchIdx = d.branch;
                if (x->w.stype == SRC_DC && d.p.size() >= 3) { x->acMag = d.p[1]; x->acPhase = d.p[2]; }
                // Value: DC -> value, SIN -> offset, PULSE -> v1, PWL -> values[0] (vsource.go:36-96)
                x->Value = (x->w.stype == SRC_PULSE) ? x->w.v1 : (x->w.stype == SRC_PWL ? x->w.values[0] : x->w.dcValue);
                dev.reset(x); break;
            }
            case K_I: {
                auto* x = new ISource; fill_waveform(x->w, d);
                if (x->w.stype == SRC_DC && d.p.size() >= 3) { x->acMag = d.p[1]; x->acPhase = d.p[2]; }
                x->Value = (x->w.stype == SRC_PULSE) ? x->w.v1 : (x->w.stype == SRC_PWL ? x->w.values[0] : x->w.dcValue);
                dev.reset(x); break;
            }
            case K_D: { auto* x = new Diode; x->Is = d.p[0]; x->N = d.p[1]; x->Tt = d.p[2]; dev.reset(x); break; }
            case K_Q: {
                auto* x = new Bjt;
                x->Ies = d.p[0]; x->Ics = d.p[1]; x->AlphaF = d.p[2]; x->Ikf = d.p[3]; x->Ikr = d.p[4];
                x->Vaf = d.p[5]; x->Var = d.p[6]; x->Nf = d.p[7]; x->Nr = d.p[8];
                x->pnp = !d.ip.empty() && d.ip[0] == 1;
                dev.reset(x); break;
            }
            case K_M: {
                auto* x = new Mosfet; const std::vector<double>& p = d.p;
                x->VTO = p[0]; x->KP = p[1]; x->GAMMA = p[2]; x->PHI = p[3]; x->LAMBDA = p[4];
                x->W = p[5]; x->L = p[6]; x->TOX = p[7]; x->CGSO = p[8]; x->CGDO = p[9]; x->CGBO = p[10];
                x->CBD = p[11]; x->CBS = p[12]; x->CJ = p[13]; x->CJSW = p[14];
                x->AS = p[15]; x->AD = p[16]; x->PS = p[17]; x->PD = p[18]; x->MJ = p[19]; x->PB = p[20];
                x->UO = p[21]; x->UCRIT = p[22]; x->UEXP = p[23]; x->VMAX = p[24];
                x->THETA = p[25]; x->ETA = p[26]; x->KAPPA = p[27]; x->DELTA = p[28];
                x->Level = d.ip.size() > 0 ? d.ip[0] : 1;
                x->pmos = d.ip.size() > 1 && d.ip[1] == 1;
                dev.reset(x); break;
            }
            case K_LCORE: {
                auto* x = new MagneticInductor; x->turns = d.p[0]; x->area = d.p[1]; x->len = d.p[2];
                x->branchIdx = d.branch; dev.reset(x); break;
            }
            case K_K: {
                auto* x = new Mutual; x->coefficient = d.p[0];
                for (int idx : d.ip) x->inductors.push_back(by_index[idx]);
                dev.reset(x); break;
            }
            default: return nullptr;
            }
            dev->name = d.name;
            for (int i = 0; i < 4; ++i) dev->n[i] = d.nodes[i];
            by_index[k] = dev.get();
            c->devices.push_back(std::move(dev));
        }
    }
    c->SetupFinish();
    return c;
}

}  // namespace

extern "C" {

struct orc_job {
    int analysis;            // 0 OP, 1 TRAN, 3 DC
    double tstart, tstop, tstep, tmax;
    int uic;
    int dc_src_dev;
    double dc_start, dc_stop, dc_inc;
    int dc2_src_dev;         // >= 0: nested sweep (dc.go:205-270), this is the inner source
    double dc2_start, dc2_stop, dc2_inc;
    int ac_sweep, ac_points; // analysis 2: 0 DEC, 1 OCT, 2 LIN; total number of points
    double ac_fstart, ac_fstop;
    int ac_refread;          // GetComplexSolution's literal index arithmetic over interleaved vectors (engine.hpp: ACAnalysis)
};

void* orc_circuit_new(int n_nodes, int n_branches) {
    Template* t = new Template; t->n_nodes = n_nodes; t->n_branches = n_branches; return t;
}
void orc_circuit_free(void* h) { delete static_cast<Template*>(h); }

int orc_circuit_add(void* h, int kind, const char* name, const int* nodes, int nn, int branch,
                    const double* p, int np, const int* ip, int nip) {
    Template* t = static_cast<Template*>(h);
    DevDesc d; d.kind = kind; d.name = name ? name : ""; d.nn = nn; d.branch = branch;
    for (int i = 0; i < nn && i < 4; ++i) d.nodes[i] = nodes[i];
    d.p.assign(p, p + np); d.ip.assign(ip, ip + nip);
    t->devs.push_back(d);
    return (int)t->devs.size() - 1;
}

// number of result columns: TRAN/DC: 1 + nodes + branches + #R ; OP: nodes + branches
int orc_n_columns(void* h, int analysis) {
    Template* t = static_cast<Template*>(h);
    if (analysis == 0) return t->n_nodes + t->n_branches;
    if (analysis == 2) { int nv = 0; for (auto& d : t->devs) if (d.kind == K_V) ++nv; return 1 + 2 * (t->n_nodes + nv); }
    int nr = 0; for (auto& d : t->devs) if (d.kind == K_R) ++nr;
    return (analysis == 4 ? 2 : 1) + t->n_nodes + t->n_branches + nr;        // 4: nested DC sweep (SWEEP1, SWEEP2)
}

// Structure introspection of one nominal build (SURVEY Appendix A checks).
//   ext2int[1..n], pivot_row[1..n], pivot_col[1..n] after the first OP factorization.
int orc_structure(void* h, int* ext2int, int* pivot_row, int* pivot_col) {
    Template* t = static_cast<Template*>(h);
    std::unique_ptr<Circuit> c = build(*t, t->devs);
    if (!c) return -1;
    int n = c->Matrix->Size;
    for (int i = 1; i <= n; ++i) ext2int[i] = c->Matrix->m.ext_to_int(i);
    OperatingPoint op; op.ckt = c.get();
    bool ok = op.Execute();
    for (int i = 1; i <= n; ++i) { pivot_row[i] = c->Matrix->m.pivot_ext_row(i); pivot_col[i] = c->Matrix->m.pivot_ext_col(i); }
    return ok ? 0 : 1;
}

// Run n_inst independent instances.  Overrides: ov_vals[k*n_inst + i] replaces parameter
// ov_par[k] of device ov_dev[k] for instance i.
//   wave     [n_inst][cap_rows][ncol] (may be NULL)    n_rows [n_inst]
//   status   [n_inst]                                   counters [n_inst][6]: accepted, rejected,
//            tran_solves, op_solves, op_path, (double bits of) fail time/value
//   stats    [n_inst][4][ncol]: min, max, sum, last over stored rows (may be NULL)
// Optional: where orc_run leaves, per instance, the signature of every pivot choice of every matrix of that instance
// (sparse13.hpp: sp13_order_sig).  Set before orc_run, reset to NULL after; not thread-safe across concurrent orc_run calls.
static uint64_t* g_order_sig = nullptr;
void orc_set_order_sig_buffer(uint64_t* buf) { g_order_sig = buf; }

int orc_run(void* h, const orc_job* job, int64_t n_inst, int n_ov, const int* ov_dev, const int* ov_par,
            const double* ov_vals, int n_threads, int64_t cap_rows, double* wave, int64_t* n_rows,
            int32_t* status, int64_t* counters, double* stats) {
    Template* t = static_cast<Template*>(h);
    const bool nested = job->analysis == 3 && job->dc2_src_dev >= 0;
    const int ncol = orc_n_columns(h, nested ? 4 : job->analysis);
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if ((int64_t)n_threads > n_inst) n_threads = (int)n_inst;
    std::atomic<int64_t> next(0);
    std::atomic<int> bad(0);
    auto worker = [&]() {
        std::vector<DevDesc> devs = t->devs;
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n_inst) break;
            for (int k = 0; k < n_ov; ++k) devs[ov_dev[k]].p[ov_par[k]] = ov_vals[(int64_t)k * n_inst + i];
            sp13_order_sig_reset();
            std::unique_ptr<Circuit> c = build(*t, devs);
            if (!c) { bad = 1; break; }
            ResultStore rs; rs.nsig = ncol;
            int st = RUN_OK; Counters cnt; int op_path = 0; double fail_at = 0;
            if (job->analysis == 0) {
                OperatingPoint op; op.ckt = c.get();
                bool ok = op.Execute();
                st = ok ? RUN_OK : RUN_OP_FAILED;
                op_path = op.path;
                cnt.op_solves = c->Matrix->n_solves;
                if (ok) rs.push(op.result.data() + 1);
            } else if (job->analysis == 2) {
                if (!c->nonlinearDevices.empty()) { bad = 3; break; }      // refused, like the product (engine.hpp: ACAnalysis)
                ACAnalysis ac(job->ac_fstart, job->ac_fstop, job->ac_points, job->ac_sweep);
                ac.ckt = c.get(); ac.interleaved_readout = job->ac_refread != 0;
                st = ac.Setup();
                const long op_solves = c->Matrix->n_solves;
                if (st == RUN_OK) st = ac.Execute(rs);
                cnt.op_solves = c->Matrix->n_solves - op_solves;           // the sweep's solves (the product runs no operating point)
                fail_at = ac.fail_freq;
            } else if (job->analysis == 1) {
                Transient tr(job->tstart, job->tstop, job->tstep, job->tmax, job->uic != 0);
                st = tr.Setup(c.get());
                if (st == RUN_OK) st = tr.Execute(rs);
                cnt = tr.cnt; op_path = tr.op_path; fail_at = tr.fail_time;
                if (st == RUN_OP_FAILED) cnt.op_solves = c->Matrix->n_solves;
            } else {
                DCSweep dc(job->dc_start, job->dc_stop, job->dc_inc);
                dc.ckt = c.get();
                Device* sd = nullptr; Device* sd2 = nullptr;
                // map template device index -> built device (K devices are moved last)
                { int pos = 0; for (size_t k = 0; k < devs.size(); ++k) { if (devs[k].kind == K_K) continue; if ((int)k == job->dc_src_dev) sd = c->devices[pos].get(); if (nested && (int)k == job->dc2_src_dev) sd2 = c->devices[pos].get(); ++pos; } }
                if (!sd || sd->type != 'V') { bad = 2; break; }
                if (nested && (!sd2 || sd2->type != 'V')) { bad = 2; break; }
                dc.source = static_cast<VSource*>(sd);
                if (nested) {
                    dc.SetSecond(static_cast<VSource*>(sd2), job->dc2_start, job->dc2_stop, job->dc2_inc);
                    st = dc.ExecuteNested(rs);
                } else st = dc.Execute(rs);
                cnt.op_solves = c->Matrix->n_solves; fail_at = dc.fail_val;
            }
            if (g_order_sig) g_order_sig[i] = sp13_order_sig;
            if (status) status[i] = st;
            if (n_rows) n_rows[i] = rs.n_rows;
            if (counters) {
                int64_t* cc = counters + i * 6;
                cc[0] = cnt.accepted; cc[1] = cnt.rejected; cc[2] = cnt.tran_solves; cc[3] = cnt.op_solves; cc[4] = op_path;
                std::memcpy(&cc[5], &fail_at, 8);
            }
            if (wave) {
                int64_t nr = rs.n_rows < cap_rows ? rs.n_rows : cap_rows;
                std::memcpy(wave + (size_t)i * cap_rows * ncol, rs.rows.data(), (size_t)nr * ncol * sizeof(double));
            }
            if (stats) {
                double* s = stats + (size_t)i * 4 * ncol;
                for (int j = 0; j < ncol; ++j) { s[j] = INFINITY; s[ncol + j] = -INFINITY; s[2 * ncol + j] = 0; s[3 * ncol + j] = 0; }
                for (long r = 0; r < rs.n_rows; ++r)
                    for (int j = 0; j < ncol; ++j) {
                        double v = rs.rows[(size_t)r * ncol + j];
                        s[j] = std::fmin(s[j], v); s[ncol + j] = std::fmax(s[ncol + j], v);
                        s[2 * ncol + j] += v; s[3 * ncol + j] = v;
                    }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int k = 1; k < n_threads; ++k) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return bad.load();
}

// Exposed helpers for unit tests.
double orc_go_sin(double x) { return go_sin(x); }
double orc_go_max(double x, double y) { return go_max(x, y); }
double orc_go_min(double x, double y) { return go_min(x, y); }
double orc_go_pow(double x, double y) { return go_pow(x, y); }
int orc_format_value_factor(double v, char* buf, int cap) {
    std::string s = format_value_factor(v);
    snprintf(buf, cap, "%s", s.c_str());
    return (int)s.size();
}
int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// The reference's matrix operator on a batch of dense systems (pkg/matrix/circuit.go:57-63 SetupElements creates every
// element; :126-150 Solve = Factor + Solve).  The first Factor() — order-and-factor — runs on A_nominal and freezes the
// order (the product's contract: the order of the NOMINAL instance); every instance is then Clear / AddElement /
// Factor (re-use) / Solve.  prow / pcol (n entries, 1-based external) return the chosen order.  Checker for
// tsb_lu_solve_batched (tests/test_lu_operator.py).
int orc_lu_batch(int n, const double* A_nominal, int64_t n_inst, const double* A, const double* b, double* x, int32_t* status,
                 int* prow, int* pcol) {
    Sparse13 m(n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m.get_element(i + 1, j + 1) += A_nominal[i * n + j];
    if (m.factor() >= SP_ZERO_DIAG) return -1;
    for (int k = 1; k <= n; ++k) { prow[k - 1] = m.pivot_ext_row(k); pcol[k - 1] = m.pivot_ext_col(k); }
    std::vector<double> rhs(n + 1), sol(n + 1);
    for (int64_t q = 0; q < n_inst; ++q) {
        m.clear();
        const double* Aq = A + q * n * n;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) m.get_element(i + 1, j + 1) += Aq[i * n + j];
        if (m.factor() >= SP_ZERO_DIAG) { status[q] = 1; for (int i = 0; i < n; ++i) x[q * n + i] = 0.0; continue; }
        rhs[0] = 0.0;
        for (int i = 0; i < n; ++i) rhs[i + 1] = b[q * n + i];
        m.solve(rhs, sol);
        for (int i = 0; i < n; ++i) x[q * n + i] = sol[i + 1];
        status[q] = 0;
    }
    return 0;
}

}  // extern "C"
