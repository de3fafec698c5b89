// ORACLE — TEST INFRASTRUCTURE ONLY (see sparse13.hpp header).  PARITY UNPINNED: the
// reference ships no tests / golden vectors and cannot be built here (no Go toolchain).
//
// engine.hpp — CPU restatement of the reference's solver path, statement by statement:
//   pkg/matrix/circuit.go   -> CircuitMatrix
//   pkg/device/*.go         -> Resistor, Capacitor, Inductor, VSource, ISource, Diode, Bjt,
//                              Mosfet, Mutual, MagneticInductor
//   pkg/circuit/circuit.go  -> Circuit
//   pkg/analysis/*.go       -> OperatingPoint, DCSweep, Transient, BaseAnalysis
//   pkg/util/formatter.go   -> format_value_factor
// Every quirk listed in SURVEY.md §3.6 (Q1-Q27) is reproduced on purpose.  Compile with
// -ffp-contract=off: the Go compiler does not fuse multiply-adds on amd64.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "gomath.hpp"
#include "sparse13.hpp"

namespace orc {

// internal/consts/consts.go:4-6
static const double CHARGE = 1.6021918e-19;
static const double BOLTZMANN = 1.3806226e-23;
static const double KELVIN = 273.15;

enum Mode { OP_MODE = 0, TRAN_MODE = 1, AC_MODE = 2, DC_MODE = 3 };   // device.go:64-71

// device.go:83-94 (only the fields that are ever read)
struct Status {
    double Time = 0, TimeStep = 0, Gmin = 0, Frequency = 0;
    int Mode = OP_MODE;
    double Temp = 0;
};

// pkg/matrix/circuit.go
struct CircuitMatrix {
    int Size;
    Sparse13 m;
    std::vector<double> rhs, sol;
    // AC analysis (a circuit built with isComplex, matrix/circuit.go:20-55): the LOGICAL complex right-hand side and solution,
    // entry i = re[i] + j*im[i].  The reference keeps them interleaved in one float64 slice (AddComplexRHS: rhs[2i], rhs[2i+1]);
    // what that does to its read-out is in ACAnalysis::Execute.
    std::vector<double> rhs_im, sol_re, sol_im;
    long n_solves = 0;
    explicit CircuitMatrix(int size) : Size(size), m(size), rhs(size + 1, 0.0), sol(size + 1, 0.0), rhs_im(size + 1, 0.0) {}
    void SetupElements() {                       // circuit.go:57-63
        for (int i = 1; i <= Size; ++i) for (int j = 1; j <= Size; ++j) m.get_element(i, j);
    }
    void AddElement(int i, int j, double v) {    // :65-71
        if (i <= 0 || j <= 0 || i > Size || j > Size) return;
        m.get_element(i, j) += v;
    }
    void AddRHS(int i, double v) {               // :99-105
        if (i <= 0 || i > Size) return;
        rhs[i] += v;
    }
    void AddComplexElement(int i, int j, double re, double im) {   // :73-83
        if (i <= 0 || j <= 0 || i > Size || j > Size) return;
        m.get_element(i, j) += re;
        m.get_element_imag(i, j) += im;
    }
    void AddComplexRHS(int i, double re, double im) {              // :85-97
        if (i <= 0 || i > Size) return;
        rhs[i] += re;
        rhs_im[i] += im;
    }
    bool SolveComplex() {                        // :126-150 with config.Complex
        ++n_solves;
        if (m.factor_complex() >= SP_ZERO_DIAG) return false;
        m.solve_complex(rhs, rhs_im, sol_re, sol_im);
        return true;
    }
    void LoadGmin(double gmin) {                 // :107-114
        for (int i = 1; i <= Size; ++i) if (double* d = m.diag(i)) *d += gmin;
    }
    void Clear() {                               // :116-124
        m.clear();
        std::fill(rhs.begin(), rhs.end(), 0.0);
        std::fill(rhs_im.begin(), rhs_im.end(), 0.0);
    }
    bool Solve() {                               // :126-150
        ++n_solves;
        int err = m.factor();
        if (err >= SP_ZERO_DIAG) return false;
        m.solve(rhs, sol);
        return true;
    }
    const std::vector<double>& Solution() const { return sol; }
};

struct Device {
    std::string name;
    char type = '?';
    int n[4] = {0, 0, 0, 0};
    double Value = 0;
    virtual ~Device() {}
    virtual bool is_nonlinear() const { return false; }       // satisfies device.NonLinear
    virtual bool is_time_dependent() const { return false; }  // satisfies device.TimeDependent
    virtual void Stamp(CircuitMatrix& mx, const Status& st) = 0;
    virtual void UpdateVoltages(const std::vector<double>&) {}
    virtual void LoadState(const std::vector<double>&, const Status&) {}
    virtual void UpdateState(const std::vector<double>&, const Status&) {}
    virtual double CalculateLTE(const Status&) { return 0; }
    virtual double GetValue() { return Value; }
    // InductorComponent (device.go:47-56)
    virtual bool is_inductor_component() const { return false; }
    virtual double GetCurrent() const { return 0; }
    virtual int BranchIndex() const { return 0; }
};

// ---------------------------------------------------------------- resistor.go
struct Resistor : Device {
    double Tc1 = 0, Tc2 = 0, Tnom = 300.15;
    Resistor() { type = 'R'; }
    double temperatureAdjustedValue(double temp) const {   // :77-81
        double dt = temp - Tnom;
        double factor = 1.0 + Tc1 * dt + Tc2 * dt * dt;
        return Value * factor;
    }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :32-75
        int n1 = n[0], n2 = n[1];
        double g = 1.0 / temperatureAdjustedValue(st.Temp);
        if (st.Mode == AC_MODE) {                          // :43-54
            if (n1 != 0) { mx.AddComplexElement(n1, n1, g, 0); if (n2 != 0) mx.AddComplexElement(n1, n2, -g, 0); }
            if (n2 != 0) { if (n1 != 0) mx.AddComplexElement(n2, n1, -g, 0); mx.AddComplexElement(n2, n2, g, 0); }
            return;
        }
        if (n1 != 0) { mx.AddElement(n1, n1, g); if (n2 != 0) mx.AddElement(n1, n2, -g); }
        if (n2 != 0) { if (n1 != 0) mx.AddElement(n2, n1, -g); mx.AddElement(n2, n2, g); }
    }
};

// ---------------------------------------------------------------- capacitor.go
struct Capacitor : Device {
    double Voltage0 = 0, Voltage1 = 0, current0 = 0, current1 = 0, charge0 = 0, charge1 = 0;
    double Tc1 = 0, Tc2 = 0, Tnom = 300.15;
    Capacitor() { type = 'C'; }
    bool is_time_dependent() const override { return true; }
    double temperatureAdjustedValue(double temp) const {   // :180-184
        double dt = temp - Tnom;
        double factor = 1.0 + Tc1 * dt + Tc2 * dt * dt;
        return Value * factor;
    }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :43-109
        int n1 = n[0], n2 = n[1];
        double adjustedC = temperatureAdjustedValue(st.Temp);
        if (st.Mode == AC_MODE) {                          // :48-66
            double omega = 2 * M_PI * st.Frequency;
            double im = omega * adjustedC;
            if (n1 != 0) { mx.AddComplexElement(n1, n1, 0.0, im); if (n2 != 0) mx.AddComplexElement(n1, n2, -0.0, -im); }
            if (n2 != 0) { mx.AddComplexElement(n2, n2, 0.0, im); if (n1 != 0) mx.AddComplexElement(n2, n1, -0.0, -im); }
        } else if (st.Mode == OP_MODE) {
            double gmin = st.Gmin;
            if (gmin < 1e-12) gmin = 1e-12;
            if (n1 != 0) { mx.AddElement(n1, n1, gmin); if (n2 != 0) mx.AddElement(n1, n2, -gmin); }
            if (n2 != 0) { mx.AddElement(n2, n2, gmin); if (n1 != 0) mx.AddElement(n2, n1, -gmin); }
        } else if (st.Mode == TRAN_MODE) {
            double dt = st.TimeStep;
            double geq = adjustedC / dt;
            double ceq = charge1 / dt;
            if (n1 != 0) {
                mx.AddElement(n1, n1, geq);
                if (n2 != 0) mx.AddElement(n1, n2, -geq);
                mx.AddRHS(n1, ceq);
            }
            if (n2 != 0) {
                mx.AddElement(n2, n2, geq);
                if (n1 != 0) mx.AddElement(n2, n1, -geq);
                mx.AddRHS(n2, -ceq);
            }
        }
        // DCSweep status uses Mode=OperatingPointAnalysis (dc.go:113-117), so no other case occurs.
    }
    static double vdiff(const int* n, const std::vector<double>& v) {
        double v1 = 0, v2 = 0;
        if (n[0] != 0) v1 = v[n[0]];
        if (n[1] != 0) v2 = v[n[1]];
        return v1 - v2;
    }
    void LoadState(const std::vector<double>& v, const Status& st) override {   // :111-124
        double vd = vdiff(n, v);
        current0 = Value * (vd - Voltage0) / st.TimeStep;
    }
    void UpdateState(const std::vector<double>& v, const Status&) override {    // :155-171
        double vd = vdiff(n, v);
        charge1 = charge0;
        charge0 = Value * vd;
        Voltage1 = Voltage0;
        Voltage0 = vd;
    }
    double CalculateLTE(const Status& st) override {                            // :173-178
        double qNew = Value * Voltage0;
        double qOld = Value * Voltage1;
        return std::fabs(qNew - qOld) / (2.0 * st.TimeStep);
    }
};

// util/integrator.go:33-48 with order 1: coeffs[0] = 1.0 / (beta * dt), beta = 1.0
static inline double bdf1_coeff0(double dt) { return 1.0 / (1.0 * dt); }

// ---------------------------------------------------------------- inductor.go
struct Inductor : Device {
    double Current0 = 0, Current1 = 0, Voltage0 = 0, Voltage1 = 0, flux0 = 0, flux1 = 0;
    int branchIdx = 0;
    Inductor() { type = 'L'; }
    bool is_time_dependent() const override { return true; }
    bool is_inductor_component() const override { return true; }
    double GetCurrent() const override { return Current0; }
    int BranchIndex() const override { return branchIdx; }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :38-79
        int n1 = n[0], n2 = n[1], b = branchIdx;
        if (st.Mode == AC_MODE) {
            // :43-57 — j*omega*L as an ADMITTANCE between the nodes; the branch row and column stay empty, so the matrix of
            // any circuit with an inductor is singular in the reference's AC analysis (kept: a quirk, like Q9)
            double omega = 2 * M_PI * st.Frequency;
            if (n1 != 0) { mx.AddComplexElement(n1, n1, 0, omega * Value); if (n2 != 0) mx.AddComplexElement(n1, n2, 0, -omega * Value); }
            if (n2 != 0) { mx.AddComplexElement(n2, n2, 0, omega * Value); if (n1 != 0) mx.AddComplexElement(n2, n1, 0, -omega * Value); }
            return;
        }
        if (n1 != 0) { mx.AddElement(n1, b, -1); mx.AddElement(b, n1, -1); }
        if (n2 != 0) { mx.AddElement(n2, b, 1); mx.AddElement(b, n2, 1); }
        double dt = st.TimeStep;
        if (dt <= 0) dt = 1e-9;
        double c0 = bdf1_coeff0(dt);
        mx.AddElement(b, b, -c0 * Value);
        mx.AddRHS(b, c0 * Value * Current1);
    }
    void LoadState(const std::vector<double>& v, const Status& st) override {   // :81-95
        double vd = Capacitor::vdiff(n, v);
        double dt = st.TimeStep;
        Current0 = Current1 + (vd * dt) / Value;
        flux0 = flux1 + vd * dt;
    }
    void UpdateState(const std::vector<double>& v, const Status&) override {    // :97-114
        double v1 = 0, v2 = 0;
        if (n[0] != 0) v1 = v[n[0]];
        if (n[1] != 0) v2 = v[n[1]];
        Voltage1 = Voltage0;
        Voltage0 = v1 - v2;
        Current1 = Current0;
        double equivR = Value / 1e-9;
        Current0 = Voltage0 / equivR;
    }
    double CalculateLTE(const Status& st) override {                            // :116-121
        double currentLTE = std::fabs(Current0 - Current1) / (2.0 * st.TimeStep);
        double voltageLTE = std::fabs(Voltage0 - Voltage1) / (2.0 * st.TimeStep);
        return go_max(currentLTE, voltageLTE);
    }
};

// ---------------------------------------------------------------- vsource.go / isource.go
enum SrcType { SRC_DC = 0, SRC_SIN = 1, SRC_PULSE = 2, SRC_PWL = 3 };

struct Waveform {
    int stype = SRC_DC;
    double dcValue = 0, amplitude = 0, freq = 0, phase = 0;
    double v1 = 0, v2 = 0, delay = 0, rise = 0, fall = 0, pWidth = 0, period = 0;
    std::vector<double> times, values;
    double eval(double t) const {                 // vsource.go:113-127, isource.go:111-125
        switch (stype) {
        case SRC_DC: return dcValue;
        case SRC_SIN: {
            double phaseRad = phase * M_PI / 180.0;
            return dcValue + amplitude * go_sin(2.0 * M_PI * freq * t + phaseRad);
        }
        case SRC_PULSE: return pulse(t);
        case SRC_PWL: return pwl(t);
        }
        return 0;
    }
    double pulse(double t) const {                // vsource.go:179-209
        if (t < delay) return v1;
        t = t - delay;
        if (period > 0) t = std::fmod(t, period);
        if (t < rise) {
            if (rise == 0) return v2;
            return v1 + (v2 - v1) * t / rise;
        }
        if (t < rise + pWidth) return v2;
        double fallStart = rise + pWidth;
        if (t < fallStart + fall) {
            if (fall == 0) return v1;
            return v2 - (v2 - v1) * (t - fallStart) / fall;
        }
        return v1;
    }
    double pwl(double t) const {                  // vsource.go:211-231
        if (t <= times[0]) return values[0];
        size_t last = times.size() - 1;
        if (t >= times[last]) return values[last];
        for (size_t i = 1; i < times.size(); ++i) {
            if (t <= times[i]) {
                double t1 = times[i - 1], t2 = times[i];
                double a = values[i - 1], b = values[i];
                double slope = (b - a) / (t2 - t1);
                return a + slope * (t - t1);
            }
        }
        return values[last];
    }
};

struct VSource : Device {
    Waveform w;
    int branchIdx = 0;
    VSource() { type = 'V'; }
    double acMag = 0, acPhase = 0;                               // NewACVoltageSource (vsource.go:98-111); 0 for every other kind
    int BranchIndex() const override { return branchIdx; }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // vsource.go:131-152
        int n1 = n[0], n2 = n[1], b = branchIdx;
        if (st.Mode == AC_MODE) {                                // StampAC, :155-177
            double phaseRad = acPhase * M_PI / 180.0;
            double voltageReal = acMag * std::cos(phaseRad);
            double voltageImag = acMag * go_sin(phaseRad);
            if (n1 != 0) { mx.AddComplexElement(b, n1, 1.0, 0.0); mx.AddComplexElement(n1, b, 1.0, 0.0); }
            if (n2 != 0) { mx.AddComplexElement(b, n2, -1.0, 0.0); mx.AddComplexElement(n2, b, -1.0, 0.0); }
            mx.AddComplexRHS(b, voltageReal, voltageImag);
            return;
        }
        if (n1 != 0) { mx.AddElement(b, n1, 1); mx.AddElement(n1, b, 1); }
        if (n2 != 0) { mx.AddElement(b, n2, -1); mx.AddElement(n2, b, -1); }
        mx.AddRHS(b, w.eval(st.Time));
    }
    void SetValue(double v) { Value = v; w.dcValue = v; }        // :241-244
};

struct ISource : Device {
    Waveform w;
    ISource() { type = 'I'; }
    double acMag = 0, acPhase = 0;
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // isource.go:130-147
        int n1 = n[0], n2 = n[1];
        if (st.Mode == AC_MODE) {                                // StampAC, :149-165
            double acPhaseRad = acPhase * M_PI / 180.0;
            double currentReal = acMag * std::cos(acPhaseRad);
            double currentImag = acMag * go_sin(acPhaseRad);
            if (n1 != 0) mx.AddComplexRHS(n1, currentReal, currentImag);
            if (n2 != 0) mx.AddComplexRHS(n2, -currentReal, -currentImag);
            return;
        }
        double current = w.eval(st.Time);
        if (n1 != 0) mx.AddRHS(n1, current);
        if (n2 != 0) mx.AddRHS(n2, -current);
    }
};

// ---------------------------------------------------------------- diode.go
struct Diode : Device {
    double Is = 1e-14, N = 1.0, Gmin = 1e-12, Eg = 1.11, Xti = 3.0, Tt = 0.0;   // :66-84
    double vd = 0, id = 0, charge = 0, gd = 0, prevCharge = 0, capCurrent = 0;
    Diode() { type = 'D'; }
    bool is_nonlinear() const override { return true; }
    // NOT TimeDependent (Q11): SetTimeStep has the wrong signature, LoadState is missing.
    static double thermalVoltage(double temp) {                  // :78-84
        if (temp <= 0) temp = 300.15;
        return BOLTZMANN * temp / CHARGE;
    }
    double temperatureAdjustedIs(double temp) const {            // :108-117
        const double ktemp = KELVIN + 27;
        double vt = thermalVoltage(temp);
        double ratio = temp / ktemp;
        double egfact = -Eg / (2 * vt) * (temp / ktemp - 1.0);
        return Is * go_pow(ratio, Xti / N) * std::exp(egfact);
    }
    double calculateCurrent(double v, double temp) const {       // :119-135
        double vt = thermalVoltage(temp);
        double nvt = N * vt;
        if (v > -3.0 * nvt) {
            double arg = v / nvt;
            if (arg > 40.0) arg = 40.0;
            double evd = std::exp(arg);
            double is_t = temperatureAdjustedIs(temp);
            return is_t * (evd - 1.0);
        }
        return -temperatureAdjustedIs(temp);
    }
    double calculateConductance(double v, double i, double temp) const {   // :137-148
        double vt = thermalVoltage(temp);
        double nvt = N * vt;
        if (v > -3.0 * nvt) return (std::fabs(i) + temperatureAdjustedIs(temp)) / nvt + Gmin;
        return Gmin;
    }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :184-227
        id = calculateCurrent(vd, st.Temp);
        gd = calculateConductance(vd, id, st.Temp);
        if (st.Mode == TRAN_MODE) {
            charge = Tt * id;
            if (st.TimeStep > 0) {
                capCurrent = (charge - prevCharge) / st.TimeStep;
                double geq = Tt * gd / st.TimeStep;
                gd += geq;
                id += capCurrent;
            }
        }
        int n1 = n[0], n2 = n[1];
        if (n1 != 0) {
            mx.AddElement(n1, n1, gd);
            if (n2 != 0) mx.AddElement(n1, n2, -gd);
            mx.AddRHS(n1, -(id - gd * vd));
        }
        if (n2 != 0) {
            if (n1 != 0) mx.AddElement(n2, n1, -gd);
            mx.AddElement(n2, n2, gd);
            mx.AddRHS(n2, (id - gd * vd));
        }
    }
    void UpdateVoltages(const std::vector<double>& v) override { // :307-324
        double v1 = 0, v2 = 0;
        if (n[0] != 0) v1 = v[n[0]];
        if (n[1] != 0) v2 = v[n[1]];
        vd = v1 - v2;
    }
};

// ---------------------------------------------------------------- bjt.go
struct Bjt : Device {
    bool pnp = false;
    double Ies = 1e-15, Ics = 1e-15, AlphaF = 0.98, Nf = 1.0, Nr = 1.0;
    double Ikf = 1e-3, Ikr = 1e-3, Vaf = 50.0, Var = 50.0;     // :93-105 (parser overrides)
    double vbe = 0, vbc = 0, vce = 0, ic = 0, ib = 0, ie = 0, gm = 0, gpi = 0, gout = 0;
    Bjt() { type = 'Q'; }
    bool is_nonlinear() const override { return true; }
    static double thermalVoltage(double temp) {                  // :122-127
        if (temp <= 0) temp = 300.15;
        return BOLTZMANN * temp / CHARGE;
    }
    void calculateInitialOperatingPoint(double temp) {           // :110-120
        double vt = thermalVoltage(temp);
        double targetIc = 1e-3;
        vbe = Nf * vt * std::log(targetIc / Ies);
        vce = go_max(2.0, vbe + 1.0);
        vbc = vbe - vce;
    }
    void calculateCurrents(double temp) {                        // :214-255
        double vt = thermalVoltage(temp);
        double expVbe = std::exp(vbe / (Nf * vt));
        double expVbc = std::exp(vbc / (Nr * vt));
        double sign = pnp ? -1.0 : 1.0;
        double iF0 = sign * Ies * (expVbe - 1);
        double iR0 = sign * Ics * (expVbc - 1);
        double iF = iF0;
        if (Vaf > 0) iF = iF0 * (1 - vbc / Vaf);
        double iR = iR0;
        if (Var > 0) iR = iR0 * (1 + vbe / Var);
        double qb = 1.0;
        if (Vaf > 0) qb = 1.0 / (1 - vbc / Vaf);
        if (Ikf > 0) iF = iF / (1 + std::fabs(iF) / (Ikf * qb));
        if (Ikr > 0) iR = iR / (1 + std::fabs(iR) / (Ikr * qb));
        double IE = sign * (iF - iR);
        double IC = sign * ((AlphaF * iF - iR) / qb);
        double IB = IE - IC;
        ie = IE; ic = IC; ib = IB;
    }
    void calculateConductances(double temp) {                    // :257-281
        double vt = thermalVoltage(temp);
        double expVbe = std::exp(vbe / (Nf * vt));
        double dIes_dVbe = Ies * expVbe / (Nf * vt);
        double qb = 1.0;
        if (Vaf > 0) qb = 1.0 / (1 - vbc / Vaf);
        gm = AlphaF * dIes_dVbe / qb;
        if (vt != 0) gpi = std::fabs(ib) / vt; else gpi = 1e-12;
        if (Vaf != 0) gout = AlphaF * Ies * (expVbe - 1) * (1 / Vaf) * go_pow(1 + vce / Vaf, -2);
        else gout = 1e-12;
    }
    void UpdateVoltages(const std::vector<double>& v) override { // :283-313
        double vc = 0, vb = 0, ve = 0;
        if (n[0] != 0) vc = v[n[0]];
        if (n[1] != 0) vb = v[n[1]];
        if (n[2] != 0) ve = v[n[2]];
        if (pnp) { vbe = ve - vb; vbc = vc - vb; vce = ve - vc; }
        else     { vbe = vb - ve; vbc = vb - vc; vce = vc - ve; }
    }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :315-374
        int nc = n[0], nb = n[1], ne = n[2];
        if (vbe == 0 && vce == 0) calculateInitialOperatingPoint(st.Temp);
        calculateCurrents(st.Temp);
        calculateConductances(st.Temp);
        // calculateCapacitances(): results (Cbe, Cbc) are only read by AC / dead code.
        if (nc != 0) {
            mx.AddElement(nc, nc, gout);
            if (nb != 0) mx.AddElement(nc, nb, -gout - gm);
            if (ne != 0) mx.AddElement(nc, ne, gm);
            mx.AddRHS(nc, -ic + gout * vce);
        }
        if (nb != 0) {
            mx.AddElement(nb, nb, gpi);
            if (nc != 0) mx.AddElement(nb, nc, -gpi);
            mx.AddRHS(nb, -ib + gpi * vbe);
        }
        if (ne != 0) {
            mx.AddElement(ne, ne, gpi + gm);
            if (nb != 0) mx.AddElement(ne, nb, -gpi - gm);
            mx.AddRHS(ne, -ie);
        }
    }
};

// ---------------------------------------------------------------- mosfet.go
struct Mosfet : Device {
    bool pmos = false;
    int Level = 1;
    double L = 10e-6, W = 10e-6, AD = 0, AS = 0, PD = 0, PS = 0;
    double VTO = 0.7, KP = 2e-5, GAMMA = 0.5, PHI = 0.6, LAMBDA = 0.01;
    double CBD = 0, CBS = 0, CGSO = 0, CGDO = 0, CGBO = 0, CJ = 0, MJ = 0.5, CJSW = 0, PB = 0.8;
    double TOX = 1e-7, UO = 600.0, UCRIT = 1e4, UEXP = 0, VMAX = 0;
    double DELTA = 0, THETA = 0, ETA = 0, KAPPA = 0.2;
    double vgs = 0, vds = 0, vbs = 0, vgd = 0, vbd = 0;
    double id = 0, gm = 0, gds = 0, gmbs = 0, cgs = 0, cgd = 0, cgb = 0;
    int region = 0;
    double qgs = 0, qgd = 0, qgb = 0, qbs = 0, qbd = 0;
    double prevQgs = 0, prevQgd = 0, prevQgb = 0, prevQbs = 0, prevQbd = 0;   // never updated (Q11)
    enum { CUTOFF = 0, LINEAR = 1, SATURATION = 2 };
    Mosfet() { type = 'M'; }
    bool is_nonlinear() const override { return true; }

    double calculateVth(double vbs_) const {                     // :296-318
        double vt0 = VTO;
        if (GAMMA > 0) {
            double vth = vt0 + GAMMA * (std::sqrt(go_max(0, PHI - vbs_)) - std::sqrt(PHI));
            if (pmos) vth = -vth;
            return vth;
        }
        if (pmos) return -vt0;
        return vt0;
    }
    void level1(double vgs_, double vds_, double vth, double& i, int& reg) const {   // :358-375
        double vgst = vgs_ - vth;
        double beta = KP * W / L;
        if (vds_ < vgst) { i = beta * (vgst * vds_ - 0.5 * vds_ * vds_) * (1.0 + LAMBDA * vds_); reg = LINEAR; }
        else { i = 0.5 * beta * vgst * vgst * (1.0 + LAMBDA * vds_); reg = SATURATION; }
    }
    void level2(double vgs_, double vds_, double vth, double& i, int& reg) const {   // :378-420
        double vgst = vgs_ - vth;
        double eps0 = 8.85e-14;
        double epsox = 3.9 * eps0;
        double cox = epsox / TOX;
        double eeff = vgst / (TOX * 100);
        double ueff = UO;
        if (UCRIT > 0 && eeff > 0) ueff /= (1.0 + go_pow(eeff / UCRIT, UEXP));
        double vdsat = vgst;
        if (VMAX > 0) {
            double ecrit = VMAX / ueff * 100;
            vdsat = go_min(vgst, ecrit * L);
        }
        double beta = ueff * cox * W / (L * 100);
        if (vds_ < vdsat) { i = beta * (vgst * vds_ - 0.5 * vds_ * vds_) * (1.0 + LAMBDA * vds_); reg = LINEAR; }
        else { i = 0.5 * beta * vdsat * vdsat * (1.0 + LAMBDA * vds_); reg = SATURATION; }
    }
    void level3(double vgs_, double vds_, double vth, double& i, int& reg) const {   // :423-459
        double vgst = vgs_ - vth;
        double vgst_eff = vgst;
        if (THETA > 0) vgst_eff = vgst / (1.0 + THETA * vgst);
        double vdsat = vgst_eff;
        if (KAPPA > 0) vdsat = vgst_eff / std::sqrt(1.0 + KAPPA * vgst_eff);
        double beta = KP * W / L;
        if (DELTA > 0) beta /= (1.0 + DELTA / W);
        if (vds_ < vdsat) {
            i = beta * (vgst_eff * vds_ - 0.5 * vds_ * vds_ / (1.0 + KAPPA * vgst_eff)) * (1.0 + LAMBDA * vds_);
            reg = LINEAR;
        } else {
            i = 0.5 * beta * vdsat * vdsat * (1.0 + LAMBDA * vds_);
            reg = SATURATION;
        }
    }
    void calculateCurrents(double vgs_, double vds_, double vbs_, double& i_out, int& reg_out) const {  // :321-355
        double sign = 1.0;
        if (pmos) { vgs_ = -vgs_; vds_ = -vds_; vbs_ = -vbs_; sign = -1.0; }
        double vth = calculateVth(vbs_);
        double vgst = vgs_ - vth;
        if (vgst <= 0) { i_out = 0.0; reg_out = CUTOFF; return; }
        double i; int reg;
        switch (Level) {
        case 2: level2(vgs_, vds_, vth, i, reg); break;
        case 3: level3(vgs_, vds_, vth, i, reg); break;
        default: level1(vgs_, vds_, vth, i, reg); break;
        }
        i_out = sign * i; reg_out = reg;
    }
    void calculateConductances() {                               // :462-537
        double sign = pmos ? -1.0 : 1.0;
        double vgs_ = vgs * sign, vds_ = vds * sign, vbs_ = vbs * sign;
        double vth = calculateVth(vbs_);
        double vgst = vgs_ - vth;
        double beta = KP * W / L;
        double gmin = 1e-12;
        if (region == CUTOFF) { gm = gmin; gds = gmin; gmbs = gmin; return; }
        if (GAMMA > 0 && PHI > 0) {
            if (vbs_ < 0) gmbs = gm * GAMMA / (2.0 * std::sqrt(PHI - vbs_));   // stale gm (Q14)
            else gmbs = gmin;
        } else gmbs = gmin;
        switch (Level) {
        case 1:
            if (region == LINEAR) {
                gm = beta * vds_ * (1.0 + LAMBDA * vds_);
                gds = beta * (vgst - vds_) * (1.0 + LAMBDA * vds_) + beta * LAMBDA * (vgst * vds_ - 0.5 * vds_ * vds_);
            } else {
                gm = beta * vgst * (1.0 + LAMBDA * vds_);
                gds = 0.5 * beta * vgst * vgst * LAMBDA;
            }
            break;
        case 2: case 3: {
            double delta = 1e-6;
            double id0 = id;
            double idg, idd, idb; int r;
            calculateCurrents(vgs_ + delta, vds_, vbs_, idg, r);
            gm = go_max((idg - id0) / delta, gmin);
            calculateCurrents(vgs_, vds_ + delta, vbs_, idd, r);
            gds = go_max((idd - id0) / delta, gmin);
            calculateCurrents(vgs_, vds_, vbs_ + delta, idb, r);
            gmbs = go_max((idb - id0) / delta, gmin);
            break;
        }
        default: break;     // other levels: Go switch has no default -> gm/gds keep stale values
        }
        gm *= sign;
        gmbs *= sign;
    }
    void calculateCapacitances() {                               // :540-594
        double cgs_ = 0, cgd_ = 0, cgb_ = 0;
        double cox = 3.9 * 8.85e-14 / TOX;
        double cgate = cox * W * L;
        double cgso = CGSO * W, cgdo = CGDO * W, cgbo = CGBO * L;
        double cbs = CBS;
        if (cbs == 0 && CJ > 0) cbs = CJ * AS + CJSW * PS;
        double cbd = CBD;
        if (cbd == 0 && CJ > 0) cbd = CJ * AD + CJSW * PD;
        CBS = cbs; CBD = cbd;
        switch (region) {
        case CUTOFF: cgb_ = 2.0 * cgate / 3.0; cgs_ = cgso; cgd_ = cgdo; break;
        case LINEAR: cgs_ = cgate / 2.0 + cgso; cgd_ = cgate / 2.0 + cgdo; cgb_ = cgbo; break;
        case SATURATION: cgs_ = 2.0 * cgate / 3.0 + cgso; cgd_ = cgdo; cgb_ = cgbo + cgate / 3.0; break;
        }
        cgs = cgs_; cgd = cgd_; cgb = cgb_;
    }
    void calculateCharges() {                                    // :597-637
        switch (region) {
        case CUTOFF: qgs = 0.0; qgd = 0.0; qgb = cgb * (vgs - vbs); break;
        default: qgs = cgs * vgs; qgd = cgd * vgd; qgb = cgb * (vgs - vbs); break;
        }
        double cbs, cbd;
        if (vbs < 0) cbs = CBS / go_pow(1.0 - vbs / PB, MJ); else cbs = CBS * (1.0 + MJ * vbs / PB);
        if (vbd < 0) cbd = CBD / go_pow(1.0 - vbd / PB, MJ); else cbd = CBD * (1.0 + MJ * vbd / PB);
        qbs = cbs * vbs;
        qbd = cbd * vbd;
    }
    void UpdateVoltages(const std::vector<double>& v) override { // :640-665 (reads v[0] for ground)
        double vg = v[n[1]], vd = v[n[0]], vs = v[n[2]], vb = v[n[3]];
        double typeValue = pmos ? -1.0 : 1.0;
        vgs = typeValue * (vg - vs);
        vds = typeValue * (vd - vs);
        vbs = typeValue * (vb - vs);
        vgd = vgs - vds;
        vbd = vbs - vds;
    }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :668-786
        int nd = n[0], ng = n[1], ns = n[2], nb = n[3];
        if (vgs == 0 && vds == 0 && vbs == 0) {
            if (!pmos) { vgs = 0.7; vds = 0.1; } else { vgs = -0.7; vds = -0.1; }
            vbs = 0.0;
            vgd = vgs - vds;
            vbd = vbs - vds;
        }
        calculateCurrents(vgs, vds, vbs, id, region);
        calculateConductances();
        calculateCapacitances();
        double gmin = st.Gmin;
        if (nd != 0) {
            mx.AddElement(nd, nd, gds + gmin);
            if (ng != 0) mx.AddElement(nd, ng, gm);
            if (ns != 0) mx.AddElement(nd, ns, -gds - gm - gmbs);
            if (nb != 0) mx.AddElement(nd, nb, gmbs);
            mx.AddRHS(nd, -id + gds * vds + gm * vgs + gmbs * vbs);
        }
        if (ns != 0) {
            mx.AddElement(ns, ns, gds + gm + gmbs + gmin);
            if (nd != 0) mx.AddElement(ns, nd, -gds);
            if (ng != 0) mx.AddElement(ns, ng, -gm);
            if (nb != 0) mx.AddElement(ns, nb, -gmbs);
            mx.AddRHS(ns, id - gds * vds - gm * vgs - gmbs * vbs);
        }
        if (st.Mode == TRAN_MODE && st.TimeStep > 0) {
            double dt = st.TimeStep;
            calculateCharges();
            double icgs = (qgs - prevQgs) / dt;
            double icgd = (qgd - prevQgd) / dt;
            double icgb = (qgb - prevQgb) / dt;
            double icbs = (qbs - prevQbs) / dt;
            double icbd = (qbd - prevQbd) / dt;
            if (ng != 0) {
                if (nd != 0) {
                    mx.AddElement(ng, nd, cgd / dt); mx.AddElement(nd, ng, cgd / dt);
                    mx.AddRHS(ng, icgd); mx.AddRHS(nd, -icgd);
                }
                if (ns != 0) {
                    mx.AddElement(ng, ns, cgs / dt); mx.AddElement(ns, ng, cgs / dt);
                    mx.AddRHS(ng, icgs); mx.AddRHS(ns, -icgs);
                }
                if (nb != 0) {
                    mx.AddElement(ng, nb, cgb / dt); mx.AddElement(nb, ng, cgb / dt);
                    mx.AddRHS(ng, icgb); mx.AddRHS(nb, -icgb);
                }
                mx.AddElement(ng, ng, (cgd + cgs + cgb) / dt);
            }
            if (nb != 0) {
                if (ns != 0) {
                    mx.AddElement(nb, ns, CBS / dt); mx.AddElement(ns, nb, CBS / dt);
                    mx.AddRHS(nb, icbs); mx.AddRHS(ns, -icbs);
                }
                if (nd != 0) {
                    mx.AddElement(nb, nd, CBD / dt); mx.AddElement(nd, nb, CBD / dt);
                    mx.AddRHS(nb, icbd); mx.AddRHS(nd, -icbd);
                }
                mx.AddElement(nb, nb, (CBD + CBS) / dt);
            }
        }
    }
};

// ---------------------------------------------------------------- magnetic.go
// 4*pi*1e-7 evaluated as an exact Go constant expression and rounded once.
static const double MU0 = 1.2566370614359172953850573533118e-6;

struct MagneticInductor : Device {
    double turns = 100, area = 1e-4, len = 0.1;
    double current0 = 0, current1 = 0;      // never advance: UpdateState is unreachable (Q11/Q12)
    int branchIdx = 0;
    MagneticInductor() { type = 'L'; }
    bool is_inductor_component() const override { return true; }
    double GetCurrent() const override { return current0; }
    int BranchIndex() const override { return branchIdx; }
    // GetValue (:147-154): core.Calculate(turns*current0/len) with dH == 0 returns dMdH = 0.
    double GetValue() override { return MU0 * (turns * turns) * area * (1 + 0.0) / len; }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :197-274
        int n1 = n[0], n2 = n[1], b = branchIdx;
        if (st.Mode == OP_MODE) {
            if (n1 != 0) { mx.AddElement(n1, b, -1); mx.AddElement(b, n1, -1); }
            if (n2 != 0) { mx.AddElement(n2, b, 1); mx.AddElement(b, n2, 1); }
            mx.AddElement(b, b, 1e-3);
            current0 = 0; current1 = 0;
        } else if (st.Mode == TRAN_MODE) {
            if (n1 != 0) { mx.AddElement(n1, b, -1); mx.AddElement(b, n1, -1); }
            if (n2 != 0) { mx.AddElement(n2, b, 1); mx.AddElement(b, n2, 1); }
            double dt = st.TimeStep;
            if (dt <= 0) dt = 1e-9;
            // `status.Time < dt || abs(current0) < 1e-9` is always true (current0 == 0):
            double L0 = MU0 * (turns * turns) * area / len;
            double diag = bdf1_coeff0(dt) * L0;
            mx.AddElement(b, b, -diag);
            mx.AddRHS(b, diag * current1);
        }
    }
};

// ---------------------------------------------------------------- mutual.go
struct Mutual : Device {
    std::vector<Device*> inductors;
    double coefficient = 0;
    Mutual() { type = 'K'; }
    void Stamp(CircuitMatrix& mx, const Status& st) override {   // :57-120
        if (st.Mode != TRAN_MODE) return;
        double dt = st.TimeStep;
        if (dt <= 0) return;
        size_t m = inductors.size();
        std::vector<int> br(m); std::vector<double> val(m), cur(m);
        for (size_t i = 0; i < m; ++i) {
            br[i] = inductors[i]->BranchIndex();
            val[i] = inductors[i]->GetValue();
            cur[i] = inductors[i]->GetCurrent();
        }
        for (size_t i = 0; i < m; ++i)
            for (size_t j = i + 1; j < m; ++j) {
                double Mij = coefficient * std::sqrt(val[i] * val[j]);
                mx.AddElement(br[i], br[j], -Mij / dt);
                mx.AddElement(br[j], br[i], -Mij / dt);
                mx.AddRHS(br[i], -Mij * cur[j] / dt);
                mx.AddRHS(br[j], -Mij * cur[i] / dt);
            }
    }
};

// ---------------------------------------------------------------- circuit.go
struct Circuit {
    int numNodes = 0, numBranches = 0;
    std::vector<std::unique_ptr<Device>> devices;     // stamp order, K last (circuit.go:83-152)
    std::vector<Device*> nonlinearDevices;
    std::vector<int> branchIdx;                       // ascending branch indices
    std::unique_ptr<CircuitMatrix> Matrix;
    Status status;                                    // *ckt.Status

    int n_signals_tran() const {
        int nr = 0;
        for (auto& d : devices) if (d->type == 'R') ++nr;
        return 1 + numNodes + numBranches + nr;
    }
    void CreateMatrix() { Matrix.reset(new CircuitMatrix(numNodes + numBranches)); }
    void SetupFinish() {                              // circuit.go:154-160
        for (auto& d : devices) if (d->is_nonlinear()) nonlinearDevices.push_back(d.get());
        Status s;                                      // CircuitStatus{Time: 0}
        Stamp(s);
        Matrix->SetupElements();
    }
    void Stamp(const Status& st) { for (auto& d : devices) d->Stamp(*Matrix, st); }
    void SetTimeStep(double dt) { status.TimeStep = dt; }   // :178-190 (device SetTimeStep only re-writes it)
    void LoadState() {                                // :192-201
        const std::vector<double>& v = Matrix->Solution();
        for (auto& d : devices) if (d->is_time_dependent()) d->LoadState(v, status);
    }
    void Update() {                                   // :203-224
        const std::vector<double>& v = Matrix->Solution();
        for (auto& d : devices) if (d->is_time_dependent()) d->UpdateState(v, status);
    }
    void UpdateNonlinearVoltages(const std::vector<double>& sol) {   // :302-313
        for (Device* d : nonlinearDevices) d->UpdateVoltages(sol);
    }
    // GetSolution (:242-273) in canonical signal order: V(node 1..), I(branch asc) = -x, I(R) in device order.
    void GetSolution(double* out) const {
        const std::vector<double>& x = Matrix->Solution();
        int k = 0;
        for (int i = 1; i <= numNodes; ++i) out[k++] = x[i];
        for (int b = numNodes + 1; b <= numNodes + numBranches; ++b) out[k++] = -x[b];
        for (auto& d : devices)
            if (d->type == 'R') {
                double v1 = 0, v2 = 0;
                if (d->n[0] > 0) v1 = x[d->n[0]];
                if (d->n[1] > 0) v2 = x[d->n[1]];
                out[k++] = (v1 - v2) / d->GetValue();
            }
    }
};

// ---------------------------------------------------------------- util/formatter.go:8-24
static inline std::string format_value_factor(double value) {
    char buf[64];
    double a = std::fabs(value);
    if (a >= 1) snprintf(buf, sizeof buf, "%.3f s", value);
    else if (a >= 1e-3) snprintf(buf, sizeof buf, "%.3f ms", value * 1e3);
    else if (a >= 1e-6) snprintf(buf, sizeof buf, "%.3f us", value * 1e6);
    else if (a >= 1e-9) snprintf(buf, sizeof buf, "%.3f ns", value * 1e9);
    else if (a >= 1e-12) snprintf(buf, sizeof buf, "%.3f ps", value * 1e12);
    else snprintf(buf, sizeof buf, "%.3e s", value);
    return buf;
}

// ---------------------------------------------------------------- analysis
struct Convergence { int maxIter = 100; double abstol = 1e-12, reltol = 1e-6, gmin = 1e-12; };   // anlysis.go:35-44

struct Counters { long accepted = 0, rejected = 0, tran_solves = 0, op_solves = 0; };

enum RunStatus { RUN_OK = 0, RUN_OP_FAILED = 1, RUN_TRAN_FAILED = 2, RUN_DC_FAILED = 3 };
enum OpPath { OP_DIRECT = 0, OP_GMIN = 1, OP_SOURCE = 2 };

struct OperatingPoint {
    Circuit* ckt = nullptr;
    Convergence conv;
    int path = OP_DIRECT;
    std::vector<double> result;   // storeResults: x[1..n] (V(node), I(branch) = +x[b]) (op.go:235-248)

    bool same(const std::vector<double>& s, const std::vector<double>& o) const {   // op.go:67-77
        for (size_t i = 1; i < s.size(); ++i) {
            double diff = std::fabs(s[i] - o[i]);
            double tol = conv.reltol * go_max(std::fabs(s[i]), std::fabs(o[i])) + conv.abstol;
            if (diff > tol) return false;
        }
        return true;
    }
    bool doNRiter(double gmin, int maxIter, const std::vector<double>* initial) {   // op.go:25-88
        CircuitMatrix& mat = *ckt->Matrix;
        std::vector<double> old = initial ? *initial : std::vector<double>(mat.Size + 1, 0.0);
        Status st; st.Time = 0; st.Mode = OP_MODE; st.Temp = 300.15; st.Gmin = gmin;
        ckt->status = st;
        for (int iter = 0; iter < maxIter; ++iter) {
            mat.Clear();
            ckt->UpdateNonlinearVoltages(old);
            ckt->Stamp(ckt->status);
            mat.LoadGmin(gmin);
            if (!mat.Solve()) return false;
            const std::vector<double>& sol = mat.Solution();
            if (iter > 0 && same(sol, old)) return true;
            old = sol;
        }
        return false;
    }
    bool calculateInitialEstimate(std::vector<double>& out) {                       // op.go:90-111
        CircuitMatrix init(ckt->Matrix->Size);
        for (auto& d : ckt->devices) if (!d->is_nonlinear()) d->Stamp(init, ckt->status);
        bool ok = init.Solve();
        ckt->Matrix->n_solves += 1;
        if (!ok) return false;
        out = init.Solution();
        return true;
    }
    bool performSourceStepping() {                                                  // op.go:113-169
        CircuitMatrix& mat = *ckt->Matrix;
        std::vector<std::pair<VSource*, double>> orig;
        for (auto& d : ckt->devices)
            if (d->type == 'V') {
                VSource* v = static_cast<VSource*>(d.get());
                orig.push_back({v, v->GetValue()});
                v->SetValue(v->GetValue() * 0.1);
            }
        struct Restore { std::vector<std::pair<VSource*, double>>& o; ~Restore() { for (auto& p : o) p.first->SetValue(p.second); } } restore{orig};
        std::vector<double> cur;
        if (!calculateInitialEstimate(cur)) cur.assign(mat.Size + 1, 0.0);
        for (double factor = 0.1; factor <= 1.0; factor += 0.1) {
            for (auto& p : orig) p.first->SetValue(p.second * factor);
            if (!doNRiter(0, conv.maxIter, &cur)) return false;
            cur = mat.Solution();
        }
        return true;
    }
    bool Execute() {                                                                // op.go:171-233
        CircuitMatrix& mat = *ckt->Matrix;
        path = OP_DIRECT;
        std::vector<double> init;
        bool have = calculateInitialEstimate(init);
        if (have) ckt->UpdateNonlinearVoltages(init);
        if (doNRiter(0, conv.maxIter, have ? &init : nullptr)) { result = mat.Solution(); return true; }
        path = OP_GMIN;
        int numGminSteps = 10;
        double startGmin = (double)mat.Size * 0.001;
        double gmin = startGmin * go_pow(10, (double)numGminSteps);
        std::vector<double> cur = mat.Solution();
        for (int i = 0; i <= numGminSteps; ++i) {
            if (!doNRiter(gmin, conv.maxIter, &cur)) break;
            cur = mat.Solution();
            gmin /= 10;
        }
        if (doNRiter(0, conv.maxIter, &cur)) { result = mat.Solution(); return true; }
        path = OP_SOURCE;
        if (!performSourceStepping()) return false;
        std::vector<double> fin = mat.Solution();
        if (!doNRiter(0, conv.maxIter, &fin)) return false;
        result = mat.Solution();
        return true;
    }
};

// Sink for stored points: row = [TIME|SWEEP1, signals...]
struct ResultStore {
    int nsig = 0;                 // including the leading TIME / SWEEP1 column
    std::vector<double> rows;     // appended rows
    long n_rows = 0;
    bool have_last = false;
    double last_time = 0;
    void push(const double* row) { rows.insert(rows.end(), row, row + nsig); ++n_rows; }
};

struct Transient {
    Circuit* ckt = nullptr;
    Convergence conv;
    OperatingPoint op;
    double time = 0, startTime, stopTime, timeStep, maxStep, minStep;
    bool useUIC;
    double trtol = 7.0;
    Counters cnt;
    double fail_time = 0;
    int op_path = 0;

    Transient(double tStart, double tStop, double tStep, double tMax, bool uic) {   // tran.go:29-55
        if (tStep > tStop / 300) tStep = tStop / 300;
        double minS = tStep / 50.0;
        if (tMax == 0) tMax = tStep;
        startTime = tStart; stopTime = tStop; timeStep = tStep; maxStep = tMax; minStep = minS; useUIC = uic;
    }
    int Setup(Circuit* c) {                                                         // tran.go:57-75
        ckt = c;
        if (!useUIC) {
            op.ckt = c;
            if (!op.Execute()) return RUN_OP_FAILED;
            op_path = op.path;
        }
        ckt->SetTimeStep(timeStep);
        return RUN_OK;
    }
    bool doNRiter(double gmin, int maxIter) {                                       // tran.go:157-216
        CircuitMatrix& mat = *ckt->Matrix;
        std::vector<double> old;
        Status st; st.Time = time; st.TimeStep = timeStep; st.Mode = TRAN_MODE; st.Temp = 300.15; st.Gmin = gmin;
        for (int iter = 0; iter < maxIter; ++iter) {
            mat.Clear();
            if (iter > 0) ckt->UpdateNonlinearVoltages(old);
            ckt->Stamp(st);
            mat.LoadGmin(gmin);
            if (!mat.Solve()) return false;
            const std::vector<double>& sol = mat.Solution();
            if (iter > 0) {
                bool all = true;
                for (size_t i = 1; i < sol.size(); ++i) {
                    double diff = std::fabs(sol[i] - old[i]);
                    double tol = conv.reltol * go_max(std::fabs(sol[i]), std::fabs(old[i])) + conv.abstol;
                    if (diff > tol) { all = false; break; }
                }
                if (all) return true;
            }
            old = sol;
        }
        return false;
    }
    double calculateTruncError() {                                                  // tran.go:239-250
        double maxLTE = 0.0;
        for (auto& d : ckt->devices)
            if (d->is_time_dependent()) {
                double lte = d->CalculateLTE(ckt->status);
                if (lte > maxLTE) maxLTE = lte;
            }
        return maxLTE;
    }
    void StoreTimeResult(ResultStore& rs, double t) {                               // anlysis.go:61-85
        if (rs.have_last) {
            if (t == rs.last_time) return;
            if (format_value_factor(t) == format_value_factor(rs.last_time)) return;
        }
        std::vector<double> row(rs.nsig);
        row[0] = t;
        ckt->GetSolution(row.data() + 1);
        rs.push(row.data());
        rs.have_last = true; rs.last_time = t;
    }
    int Execute(ResultStore& rs) {                                                  // tran.go:77-155
        if (!useUIC) {
            op.ckt = ckt;
            if (!op.Execute()) return RUN_OP_FAILED;
            op_path = op.path > op_path ? op.path : op_path;
        }
        cnt.op_solves = ckt->Matrix->n_solves;
        timeStep = minStep;
        while (time < stopTime) {
            double nextTime = time + timeStep;
            if (nextTime > stopTime) { nextTime = stopTime; timeStep = nextTime - time; }
            Status st; st.Time = time; st.TimeStep = timeStep; st.Mode = TRAN_MODE; st.Temp = 300.15; st.Gmin = conv.gmin;
            ckt->status = st;
            if (!doNRiter(0, conv.maxIter)) {
                if (timeStep > minStep) { timeStep /= 2; ++cnt.rejected; continue; }
                fail_time = time;
                cnt.tran_solves = ckt->Matrix->n_solves - cnt.op_solves;
                return RUN_TRAN_FAILED;
            }
            double lte = calculateTruncError();
            if (lte > trtol) {
                if (timeStep > minStep) { timeStep /= 2; ++cnt.rejected; continue; }
            }
            ckt->LoadState();
            ckt->Update();
            time = nextTime;
            ++cnt.accepted;
            if (time >= startTime) StoreTimeResult(rs, time);
            if (time < stopTime && timeStep < maxStep) {
                if (lte < trtol / 100) timeStep = go_min(timeStep * 2, maxStep);
                else timeStep = go_min(timeStep * 1.1, maxStep);
            }
        }
        cnt.tran_solves = ckt->Matrix->n_solves - cnt.op_solves;
        return RUN_OK;
    }
};

struct DCSweep {
    Circuit* ckt = nullptr;
    Convergence conv;
    VSource* source = nullptr;
    std::vector<double> sweepVals;
    double origVal = 0;
    double fail_val = 0;
    DCSweep(double start, double stop, double inc) {                                // dc.go:36-42
        for (double v = start; v <= stop; v += inc) sweepVals.push_back(v);
    }
    bool CheckConvergence(const std::vector<double>& o, const std::vector<double>& s) const {   // anlysis.go:46-59
        for (size_t i = 0; i < o.size(); ++i) {
            double diff = std::fabs(s[i] - o[i]);
            if (diff > conv.abstol && diff > conv.reltol * std::fabs(s[i])) return false;
        }
        return true;
    }
    bool doNRiter(double gmin, int maxIter) {                                       // dc.go:142-187
        CircuitMatrix& mat = *ckt->Matrix;
        std::vector<double> old;
        Status st; st.Mode = OP_MODE; st.Temp = 300.15; st.Gmin = gmin;
        for (int iter = 0; iter < maxIter; ++iter) {
            mat.Clear();
            if (iter > 0) ckt->UpdateNonlinearVoltages(old);
            ckt->Stamp(st);
            mat.LoadGmin(gmin);
            if (!mat.Solve()) return false;
            const std::vector<double>& sol = mat.Solution();
            if (iter > 0 && CheckConvergence(old, sol)) return true;
            old = sol;
        }
        return false;
    }
    int Execute(ResultStore& rs) {                                                  // dc.go:88-140
        origVal = source->GetValue();
        for (double val : sweepVals) {
            source->SetValue(val);
            Status st; st.Mode = OP_MODE; st.Temp = 300.15; st.Gmin = conv.gmin;
            ckt->Matrix->Clear();
            ckt->Stamp(st);                       // the "wasted" stamp (side effects on device state kept)
            if (!doNRiter(0, conv.maxIter)) { fail_val = val; return RUN_DC_FAILED; }
            std::vector<double> row(rs.nsig);
            row[0] = val;
            ckt->GetSolution(row.data() + 1);
            rs.push(row.data());
        }
        source->SetValue(origVal);
        return RUN_OK;
    }
    // nestedSweep (dc.go:205-270) + StoreNestedResult (dc.go:272-288): rows are [SWEEP1, SWEEP2, signals...]
    VSource* source2 = nullptr;
    std::vector<double> sweepVals2;
    double fail_val2 = 0;
    void SetSecond(VSource* s2, double start, double stop, double inc) {
        source2 = s2;
        for (double v = start; v <= stop; v += inc) sweepVals2.push_back(v);
    }
    int ExecuteNested(ResultStore& rs) {
        origVal = source->GetValue();
        const double origVal2 = source2->GetValue();
        for (double val1 : sweepVals) {
            source->SetValue(val1);
            for (double val2 : sweepVals2) {
                source2->SetValue(val2);
                Status st; st.Mode = OP_MODE; st.Temp = 300.15; st.Gmin = conv.gmin;
                ckt->Matrix->Clear();
                ckt->Stamp(st);
                if (!doNRiter(0, conv.maxIter)) { fail_val = val1; fail_val2 = val2; return RUN_DC_FAILED; }
                std::vector<double> row(rs.nsig);
                row[0] = val1;
                row[1] = val2;
                ckt->GetSolution(row.data() + 2);
                rs.push(row.data());
            }
        }
        source->SetValue(origVal);
        source2->SetValue(origVal2);
        return RUN_OK;
    }
};

// ---------------------------------------------------------------- ac.go
// ACAnalysis of a circuit WITHOUT nonlinear devices.  What the reference does with one (ac.go:33-98):
//   Setup    the circuit's matrix is complex from the start (cmd/spice/main.go:380-381); the operating point runs on it.  Its
//            values are real, so the first FactorComplex chooses the same pivots as a real matrix would (|re| + |im| = |re|)
//            and the order is frozen there.  The operating point's OWN numbers are scrambled — its real-indexed right-hand
//            side rhs[i] (AddRHS) is read by SolveComplex as interleaved re/im pairs — but a linear circuit converges on
//            whatever comes out (the second iteration repeats the first bit for bit), and no device of a linear circuit
//            carries operating-point state into StampAC.  So for linear circuits the sweep is independent of that artefact;
//            for nonlinear ones it is not (their gd / gm come from the scrambled point, and Bjt.Stamp never dispatches to
//            StampAC at all): those are refused, here and in the product.
//   Execute  per frequency Clear / Stamp(Mode: ACAnalysis) / Solve.  Devices whose Stamp has no AC case stamp NOTHING:
//            Mutual (mutual.go:63-65 "only for transient") and MagneticInductor (magnetic.go:205-273 switch without an AC
//            case) — their StampAC methods exist but circuit.Stamp never calls them (circuit.go:165-176).  An inductor
//            stamps j*omega*L as an admittance and leaves its branch row empty.  Either way the matrix is singular:
//            "matrix solve error at f=...".
//   Read-out GetComplexSolution(i) returns (solution[i], solution[i+Size]) (matrix/circuit.go:168-173) although the vectors are
//            interleaved (SeparatedComplexVectors: false, :41-44).  interleaved_readout = true restates that literally under
//            Sparse 1.3's layout (solution[2i] = re x_i, solution[2i+1] = im x_i, entries 0 and 1 unused = 0); false reads
//            entry i as (re x_i, im x_i) — what the accessor means if the un-vendored module returns split halves.  The module
//            is not inspectable, so both are offered; the product's default is the second.
enum { RUN_AC_FAILED = 5 };
struct ACAnalysis {
    Circuit* ckt = nullptr;
    double startFreq, stopFreq;
    int numPoints, pType;                   // 0 DEC, 1 OCT, 2 LIN
    bool interleaved_readout = false;
    std::vector<double> frequencies;
    double fail_freq = 0;
    ACAnalysis(double fStart, double fStop, int nPoints, int ptype) : startFreq(fStart), stopFreq(fStop), numPoints(nPoints), pType(ptype) {}
    void generateFrequencyPoints() {        // ac.go:100-126
        frequencies.assign(numPoints, 0.0);
        if (pType == 0) {
            double logStart = go_log10(startFreq), logStop = go_log10(stopFreq);
            double step = (logStop - logStart) / double(numPoints - 1);
            for (int i = 0; i < numPoints; ++i) frequencies[i] = go_pow(10, logStart + double(i) * step);
        } else if (pType == 1) {
            double logStart = go_log2(startFreq), logStop = go_log2(stopFreq);
            double step = (logStop - logStart) / double(numPoints - 1);
            for (int i = 0; i < numPoints; ++i) frequencies[i] = go_pow(2, logStart + double(i) * step);
        } else {
            double step = (stopFreq - startFreq) / double(numPoints - 1);
            for (int i = 0; i < numPoints; ++i) frequencies[i] = startFreq + double(i) * step;
        }
    }
    int n_signals() const {
        int nv = 0;
        for (auto& d : ckt->devices) if (d->type == 'V') ++nv;
        return 1 + 2 * (ckt->numNodes + nv);
    }
    int Setup() {                           // ac.go:33-49
        OperatingPoint op; op.ckt = ckt;
        if (!op.Execute()) return RUN_OP_FAILED;
        generateFrequencyPoints();
        return RUN_OK;
    }
    int Execute(ResultStore& rs) {          // ac.go:51-98, StoreACResult anlysis.go:87-111
        CircuitMatrix& mat = *ckt->Matrix;
        const int n = mat.Size;
        for (double freq : frequencies) {
            Status st; st.Frequency = freq; st.Mode = AC_MODE; st.Temp = 300.15;
            mat.Clear();
            ckt->Stamp(st);
            if (!mat.SolveComplex()) { fail_freq = freq; return RUN_AC_FAILED; }
            std::vector<double> flat(2 * (n + 1) + n + 2, 0.0);
            for (int i = 1; i <= n; ++i) { flat[2 * i] = mat.sol_re[i]; flat[2 * i + 1] = mat.sol_im[i]; }
            auto get = [&](int i, double& re, double& im) {
                if (interleaved_readout) { re = flat[i]; im = flat[i + n]; }
                else { re = mat.sol_re[i]; im = mat.sol_im[i]; }
            };
            std::vector<double> row(rs.nsig);
            int k = 0;
            row[k++] = freq;
            auto put = [&](int idx) {
                double re, im; get(idx, re, im);
                row[k++] = go_hypot(re, im);                            // cmplx.Abs = math.Hypot
                row[k++] = std::atan2(im, re) * 180.0 / M_PI;           // cmplx.Phase * 180 / math.Pi
            };
            for (int i = 1; i <= ckt->numNodes; ++i) put(i);
            for (auto& d : ckt->devices) if (d->type == 'V') put(d->BranchIndex());
            rs.push(row.data());
        }
        return RUN_OK;
    }
};

}  // namespace orc
