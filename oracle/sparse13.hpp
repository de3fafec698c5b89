// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.
//
// sparse13.hpp — CPU restatement of the LU arithmetic the reference obtains from the
// un-vendored Go module github.com/edp1096/sparse v0.0.0-20250223074749-e82e4651f4d6
// (reference go.mod:5), a Go port of K. Kundert's Sparse 1.3.  The module source is NOT
// under /root/reference, so this file restates the *published* Sparse 1.3 algorithm
// (spBuild.c / spFactor.c / spSolve.c, default spConfig.h: MODIFIED_MARKOWITZ off,
// DIAGONAL_PIVOTING on, DEFAULT_THRESHOLD 1e-3, TIES_MULTIPLIER 5) and is anchored on the
// reference's own call sites:
//   Create(size, cfg{Real,Translate,Expandable,ModifiedNodal,TiesMultiplier:5})  pkg/matrix/circuit.go:20-33
//   GetElement(i,j).Real += v                                                   pkg/matrix/circuit.go:65-71
//   Clear()                                                                     pkg/matrix/circuit.go:116-124
//   Factor(); Solve(rhs)                                                        pkg/matrix/circuit.go:126-150
//   Diags[i]                                                                    pkg/matrix/circuit.go:152-158
// PARITY UNPINNED: the reference ships no test or golden vector at this boundary.
//
// Representation: the orthogonal linked lists of Sparse are replaced by a dense value
// array plus an "element exists" flag per (row, col); list traversal order (ascending
// row within a column, ascending column within a row) is reproduced by index scans, so
// every floating-point operation happens in the same order as in Sparse 1.3.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <utility>

namespace orc {

// Signature of every pivot choice this thread's matrices have made since it was last reset (test aid: which instances of a
// parameter sweep does the reference order differently from the nominal one?  Their results differ from a frozen-order
// elimination by the rounding of another operation order).
inline thread_local uint64_t sp13_order_sig = 1469598103934665603ULL;
inline void sp13_order_sig_reset() { sp13_order_sig = 1469598103934665603ULL; }

enum { SP_OKAY = 0, SP_SMALL_PIVOT = 1, SP_ZERO_DIAG = 2, SP_SINGULAR = 3 };

class Sparse13 {
public:
    explicit Sparse13(int size) { create(size); }

    int size() const { return n_; }

    // spGetElement with TRANSLATE: external (row, col) -> element reference, created on demand.
    double& get_element(int ext_row, int ext_col) {
        int r, c;
        translate(ext_row, ext_col, r, c);
        char& e = ex_[idx(r, c)];
        if (!e) {
            e = 1;
            // spcCreateElement: a new element in a matrix whose rows are already linked
            // (i.e. that has been through a factorization) forces a re-ordering.
            if (rows_linked_) needs_ordering_ = true;
        }
        return v_[idx(r, c)];
    }

    // element.Imag of the same element (matrix/circuit.go:73-83 AddComplexElement): only the AC analysis writes it
    double& get_element_imag(int ext_row, int ext_col) {
        get_element(ext_row, ext_col);
        int r, c;
        translate(ext_row, ext_col, r, c);
        if (vi_.empty()) vi_.assign(v_.size(), 0.0);
        return vi_[idx(r, c)];
    }

    // spClear: zero all values, keep structure and ordering.
    void clear() {
        std::fill(v_.begin(), v_.end(), 0.0);
        std::fill(vi_.begin(), vi_.end(), 0.0);
        factored_ = false;
        error_ = SP_OKAY;
    }

    // Diag[i] in internal order (pivot positions once reordered).  nullptr if absent.
    double* diag(int i) { return ex_[idx(i, i)] ? &v_[idx(i, i)] : nullptr; }

    // spFactor.  Returns SP_OKAY / SP_SMALL_PIVOT (warnings) or >= SP_ZERO_DIAG (fatal).
    int factor() {
        if (needs_ordering_) return order_and_factor();
        // Reuse the frozen pivot order: left-looking, column by column (direct addressing
        // scatter/gather of spFactor.c; the indirect variant performs the same arithmetic).
        if (v_[idx(1, 1)] == 0.0 || !ex_[idx(1, 1)]) return zero_pivot(1);
        v_[idx(1, 1)] = 1.0 / v_[idx(1, 1)];
        std::vector<double>& dest = tmp_;
        for (int step = 2; step <= n_; ++step) {
            for (int r = 1; r <= n_; ++r) if (ex_[idx(r, step)]) dest[r] = v_[idx(r, step)];
            for (int r = 1; r < step; ++r) {
                if (!ex_[idx(r, step)]) continue;
                double u = dest[r] * v_[idx(r, r)];     // * reciprocal pivot
                v_[idx(r, step)] = u;
                for (int l = r + 1; l <= n_; ++l)
                    if (ex_[idx(l, r)]) dest[l] -= u * v_[idx(l, r)];
            }
            for (int r = step + 1; r <= n_; ++r) if (ex_[idx(r, step)]) v_[idx(r, step)] = dest[r];
            if (!ex_[idx(step, step)] || dest[step] == 0.0) return zero_pivot(step);
            v_[idx(step, step)] = 1.0 / dest[step];
        }
        factored_ = true;
        return (error_ = SP_OKAY);
    }

    // spSolve: rhs and solution are 1-based external vectors of length n+1; solution[0] = 0.
    void solve(const std::vector<double>& rhs, std::vector<double>& sol) {
        std::vector<double>& c = tmp_;
        for (int i = n_; i > 0; --i) c[i] = rhs[i2e_row_[i]];
        for (int i = 1; i <= n_; ++i) {
            double t = c[i];
            if (t != 0.0) {
                c[i] = (t *= v_[idx(i, i)]);
                for (int r = i + 1; r <= n_; ++r)
                    if (ex_[idx(r, i)]) c[r] -= t * v_[idx(r, i)];
            }
        }
        for (int i = n_; i > 0; --i) {
            double t = c[i];
            for (int col = i + 1; col <= n_; ++col)
                if (ex_[idx(i, col)]) t -= v_[idx(i, col)] * c[col];
            c[i] = t;
        }
        sol.assign(n_ + 1, 0.0);
        for (int i = n_; i > 0; --i) sol[i2e_col_[i]] = c[i];
    }

    // ---- complex matrices (AC analysis; matrix/circuit.go:130,140 FactorComplex / SolveComplex) -----------------------
    // spFactor of a complex matrix whose pivot order exists already (FactorComplexMatrix, direct-addressing form: column by
    // column, Mult = Dest[row] * (1/pivot), Dest[below] -= Mult * L; the pivot is stored as its reciprocal, formed by
    // CMPLX_RECIPROCAL = Smith's scaled division).  In the reference the order always exists when the AC sweep starts: the
    // operating point of ACAnalysis.Setup (ac.go:33-49) went through the same matrix first.
    static void crecip(double& re, double& im) {
        if ((re >= im && re > -im) || (re < im && re <= -im)) {
            double r = im / re;
            double t = 1.0 / (re + r * im);
            re = t; im = -r * t;
        } else {
            double r = re / im;
            double t = -1.0 / (im + r * re);
            im = t; re = -r * t;
        }
    }
    int factor_complex() {
        if (needs_ordering_) return (error_ = SP_SINGULAR);          // not reachable from the analyses (see above)
        if (vi_.empty()) vi_.assign(v_.size(), 0.0);
        if (!ex_[idx(1, 1)] || (std::fabs(v_[idx(1, 1)]) + std::fabs(vi_[idx(1, 1)]) == 0.0)) return zero_pivot(1);
        crecip(v_[idx(1, 1)], vi_[idx(1, 1)]);
        std::vector<double> dr(n_ + 2, 0.0), di(n_ + 2, 0.0);
        for (int step = 2; step <= n_; ++step) {
            for (int r = 1; r <= n_; ++r) if (ex_[idx(r, step)]) { dr[r] = v_[idx(r, step)]; di[r] = vi_[idx(r, step)]; }
            for (int r = 1; r < step; ++r) {
                if (!ex_[idx(r, step)]) continue;
                const double pr = v_[idx(r, r)], pi = vi_[idx(r, r)];               // reciprocal pivot
                const double mr = dr[r] * pr - di[r] * pi, mi = dr[r] * pi + di[r] * pr;   // CMPLX_MULT(Mult, Dest, *pPivot)
                v_[idx(r, step)] = mr; vi_[idx(r, step)] = mi;
                for (int l = r + 1; l <= n_; ++l)
                    if (ex_[idx(l, r)]) {                                               // CMPLX_MULT_SUBT_ASSIGN(Dest, Mult, *pElement)
                        const double er = v_[idx(l, r)], ei = vi_[idx(l, r)];
                        dr[l] -= mr * er - mi * ei;
                        di[l] -= mr * ei + mi * er;
                    }
            }
            for (int r = step + 1; r <= n_; ++r) if (ex_[idx(r, step)]) { v_[idx(r, step)] = dr[r]; vi_[idx(r, step)] = di[r]; }
            if (!ex_[idx(step, step)] || (std::fabs(dr[step]) + std::fabs(di[step]) == 0.0)) return zero_pivot(step);
            double pr = dr[step], pi = di[step];
            crecip(pr, pi);
            v_[idx(step, step)] = pr; vi_[idx(step, step)] = pi;
        }
        factored_ = true;
        return (error_ = SP_OKAY);
    }
    // SolveComplexMatrix: logical vectors (entry i = re[i] + j*im[i], 1-based external); how the reference lays them out in
    // its float64 slices is the caller's concern (engine.hpp: ACAnalysis).
    void solve_complex(const std::vector<double>& rre, const std::vector<double>& rim, std::vector<double>& sre, std::vector<double>& sim) {
        std::vector<double> cr(n_ + 2, 0.0), ci(n_ + 2, 0.0);
        for (int i = n_; i > 0; --i) { cr[i] = rre[i2e_row_[i]]; ci[i] = rim[i2e_row_[i]]; }
        for (int i = 1; i <= n_; ++i) {
            double tr = cr[i], ti = ci[i];
            if (tr != 0.0 || ti != 0.0) {
                const double pr = v_[idx(i, i)], pi = vi_[idx(i, i)];
                const double mr = tr * pr - ti * pi, mi = tr * pi + ti * pr;           // CMPLX_MULT_ASSIGN(Temp, *pPivot)
                cr[i] = mr; ci[i] = mi;
                for (int r = i + 1; r <= n_; ++r)
                    if (ex_[idx(r, i)]) {                                              // CMPLX_MULT_SUBT_ASSIGN(Intermediate[row], Temp, *pElement)
                        const double er = v_[idx(r, i)], ei = vi_[idx(r, i)];
                        cr[r] -= mr * er - mi * ei;
                        ci[r] -= mr * ei + mi * er;
                    }
            }
        }
        for (int i = n_; i > 0; --i) {
            double tr = cr[i], ti = ci[i];
            for (int col = i + 1; col <= n_; ++col)
                if (ex_[idx(i, col)]) {                                                // CMPLX_MULT_SUBT_ASSIGN(Temp, *pElement, Intermediate[col])
                    const double er = v_[idx(i, col)], ei = vi_[idx(i, col)];
                    tr -= er * cr[col] - ei * ci[col];
                    ti -= er * ci[col] + ei * cr[col];
                }
            cr[i] = tr; ci[i] = ti;
        }
        sre.assign(n_ + 1, 0.0); sim.assign(n_ + 1, 0.0);
        for (int i = n_; i > 0; --i) { sre[i2e_col_[i]] = cr[i]; sim[i2e_col_[i]] = ci[i]; }
    }

    // Introspection for the structure tests.
    int ext_to_int(int ext) const { return e2i_row_[ext]; }
    int pivot_ext_row(int step) const { return i2e_row_[step]; }
    int pivot_ext_col(int step) const { return i2e_col_[step]; }
    bool needs_ordering() const { return needs_ordering_; }
    int last_error() const { return error_; }
    int singular_step() const { return sing_step_; }

private:
    int n_ = 0, current_size_ = 0;
    std::vector<double> v_, vi_, tmp_;
    std::vector<char> ex_;
    std::vector<int> e2i_row_, e2i_col_, i2e_row_, i2e_col_;
    std::vector<long> mrow_, mcol_, mprod_;
    int singletons_ = 0;
    bool needs_ordering_ = true, factored_ = false, rows_linked_ = false;
    int error_ = SP_OKAY, sing_step_ = 0;
    double rel_threshold_ = 1.0e-3, abs_threshold_ = 0.0;
    static constexpr long TIES_MULTIPLIER = 5;          // pkg/matrix/circuit.go:28
    static constexpr long LARGEST_LONG = 0x7fffffffL;

    size_t idx(int r, int c) const { return (size_t)r * (n_ + 1) + c; }

    void create(int size) {
        n_ = size;
        v_.assign((size_t)(n_ + 1) * (n_ + 1), 0.0);
        ex_.assign(v_.size(), 0);
        tmp_.assign(n_ + 2, 0.0);
        e2i_row_.assign(n_ + 1, -1); e2i_col_.assign(n_ + 1, -1);
        i2e_row_.resize(n_ + 1); i2e_col_.resize(n_ + 1);
        for (int i = 0; i <= n_; ++i) { i2e_row_[i] = i; i2e_col_[i] = i; }
        e2i_row_[0] = e2i_col_[0] = 0;
        mrow_.assign(n_ + 2, 0); mcol_.assign(n_ + 2, 0); mprod_.assign(n_ + 3, 0);
    }

    // spBuild.c Translate(): internal numbers are handed out at first touch, row before column.
    void translate(int ext_row, int ext_col, int& r, int& c) {
        if ((r = e2i_row_[ext_row]) == -1) {
            e2i_row_[ext_row] = ++current_size_;
            e2i_col_[ext_row] = current_size_;
            r = current_size_;
            i2e_row_[r] = ext_row; i2e_col_[r] = ext_row;
        }
        if ((c = e2i_col_[ext_col]) == -1) {
            e2i_row_[ext_col] = ++current_size_;
            e2i_col_[ext_col] = current_size_;
            c = current_size_;
            i2e_row_[c] = ext_col; i2e_col_[c] = ext_col;
        }
    }

    int zero_pivot(int step) { error_ = SP_ZERO_DIAG; sing_step_ = step; return error_; }
    int matrix_is_singular(int step) { error_ = SP_SINGULAR; sing_step_ = step; return error_; }

    static double mag(double x) { return std::fabs(x); }

    // ---- spOrderAndFactor (first factorization: choose pivots, then eliminate) ----
    int order_and_factor() {
        error_ = SP_OKAY;
        rows_linked_ = true;
        int step = 1;
        count_markowitz(step);
        markowitz_products(step);
        for (; step <= n_; ++step) {
            int pr = 0, pc = 0;
            if (!search_for_pivot(step, pr, pc)) return matrix_is_singular(step);
            sp13_order_sig = (sp13_order_sig ^ (uint64_t)(((unsigned)step << 16) ^ ((unsigned)pr << 8) ^ (unsigned)pc)) * 1099511628211ULL;
            exchange_rows_and_cols(pr, pc, step);
            if (!real_row_col_elimination(step)) return error_;
            update_markowitz_numbers(step);
        }
        needs_ordering_ = false;
        factored_ = true;
        return error_;
    }

    void count_markowitz(int step) {
        for (int i = step; i <= n_; ++i) {
            long cnt = -1;
            for (int c = step; c <= n_; ++c) if (ex_[idx(i, c)]) ++cnt;
            mrow_[i] = cnt;
        }
        for (int i = step; i <= n_; ++i) {
            long cnt = -1;
            for (int r = step; r <= n_; ++r) if (ex_[idx(r, i)]) ++cnt;
            mcol_[i] = cnt;
        }
    }

    void markowitz_products(int step) {
        singletons_ = 0;
        for (int i = step; i <= n_; ++i) {
            long p = mrow_[i] * mcol_[i];
            if ((mprod_[i] = p) == 0) ++singletons_;
        }
    }

    double find_largest_in_col_from(int col, int from_row) const {
        double largest = 0.0;
        for (int r = from_row; r <= n_; ++r)
            if (ex_[idx(r, col)]) { double m = mag(v_[idx(r, col)]); if (m > largest) largest = m; }
        return largest;
    }

    // FindBiggestInColExclude: largest magnitude in column `col`, rows >= step, excluding `row`.
    double find_biggest_in_col_exclude(int row, int col, int step) const {
        double largest = 0.0;
        for (int r = step; r <= n_; ++r) {
            if (!ex_[idx(r, col)] || r == row) continue;
            double m = mag(v_[idx(r, col)]);
            if (m > largest) largest = m;
        }
        return largest;
    }

    bool search_for_pivot(int step, int& pr, int& pc) {
        if (singletons_) {
            if (search_for_singleton(step, pr, pc)) return true;
        }
        if (quickly_search_diagonal(step, pr, pc)) return true;
        if (search_diagonal(step, pr, pc)) return true;
        return search_entire_matrix(step, pr, pc);
    }

    bool acceptable(int r, int c, int step) const {
        double m = mag(v_[idx(r, c)]);
        return m > abs_threshold_ && m > rel_threshold_ * find_biggest_in_col_exclude(r, c, step);
    }

    // SearchForSingleton: scan Markowitz products from Size down to Step, Diag[Step] first.
    bool search_for_singleton(int step, int& pr, int& pc) {
        mprod_[n_ + 1] = mprod_[step];
        int singletons = singletons_--;
        mprod_[step - 1] = 0;
        int p = n_ + 1;                        // scanning position in mprod_
        while (singletons-- > 0) {
            while (mprod_[p--] != 0) { /* just passing through */ }
            int i = p + 1;
            if (i < step) break;
            if (i > n_) i = step;
            if (ex_[idx(i, i)]) {
                if (acceptable(i, i, step)) { pr = i; pc = i; return true; }
            } else {
                // Singleton off the diagonal (Sparse 1.4 form of the test; 1.3 has `!= NULL`
                // here, which makes the branch unreachable).
                if (mcol_[i] == 0) {
                    int r = step; while (r <= n_ && !ex_[idx(r, i)]) ++r;
                    if (r > n_) break;
                    if (acceptable(r, i, step)) { pr = r; pc = i; return true; }
                    if (mrow_[i] == 0) {
                        int c = step; while (c <= n_ && !ex_[idx(i, c)]) ++c;
                        if (c > n_) break;
                        if (acceptable(i, c, step)) { pr = i; pc = c; return true; }
                    }
                } else {
                    int c = step; while (c <= n_ && !ex_[idx(i, c)]) ++c;
                    if (c > n_) break;
                    if (acceptable(i, c, step)) { pr = i; pc = c; return true; }
                }
            }
        }
        singletons_++;
        return false;
    }

    // QuicklySearchDiagonal (MODIFIED_MARKOWITZ off).
    bool quickly_search_diagonal(int step, int& pr, int& pc) {
        int chosen = 0;
        long min_prod = LARGEST_LONG;
        mprod_[n_ + 1] = mprod_[step];
        mprod_[step - 1] = -1;
        int p = n_ + 2;
        for (;;) {
            while (mprod_[--p] >= min_prod) { /* just passing through */ }
            int i = p;
            if (i < step) break;
            if (i > n_) i = step;
            if (!ex_[idx(i, i)]) continue;
            double magnitude = mag(v_[idx(i, i)]);
            if (magnitude <= abs_threshold_) continue;
            if (mprod_[p] == 1) {
                // exactly one off-diagonal in the row and one in the column (reduced submatrix)
                int orow_col = 0, ocol_row = 0;
                for (int c = i + 1; c <= n_; ++c) if (ex_[idx(i, c)]) { orow_col = c; break; }
                for (int r = i + 1; r <= n_; ++r) if (ex_[idx(r, i)]) { ocol_row = r; break; }
                if (orow_col == 0 && ocol_row == 0) {
                    for (int c = step; c <= n_; ++c) if (ex_[idx(i, c)] && c != i) { orow_col = c; break; }
                    for (int r = step; r <= n_; ++r) if (ex_[idx(r, i)] && r != i) { ocol_row = r; break; }
                }
                if (orow_col != 0 && ocol_row != 0 && orow_col == ocol_row) {
                    double lo = std::fmax(mag(v_[idx(i, orow_col)]), mag(v_[idx(ocol_row, i)]));
                    if (magnitude >= lo) { pr = pc = i; return true; }
                }
            }
            min_prod = mprod_[p];
            chosen = i;
        }
        if (chosen) {
            double largest = find_biggest_in_col_exclude(chosen, chosen, step);
            if (mag(v_[idx(chosen, chosen)]) <= rel_threshold_ * largest) chosen = 0;
        }
        if (!chosen) return false;
        pr = pc = chosen;
        return true;
    }

    // SearchDiagonal: every candidate checked numerically, ties resolved by column ratio.
    bool search_diagonal(int step, int& pr, int& pc) {
        int chosen = 0;
        long min_prod = LARGEST_LONG;
        double ratio_of_accepted = 0.0;
        long ties = 0;
        mprod_[n_ + 1] = mprod_[step];
        int p = n_ + 2;
        for (int j = n_ + 1; j > step; --j) {
            --p;
            if (mprod_[p] > min_prod) continue;
            int i = (j > n_) ? step : j;
            if (!ex_[idx(i, i)]) continue;
            double magnitude = mag(v_[idx(i, i)]);
            if (magnitude <= abs_threshold_) continue;
            double largest = find_biggest_in_col_exclude(i, i, step);
            if (magnitude <= rel_threshold_ * largest) continue;
            if (mprod_[p] < min_prod) {
                chosen = i; min_prod = mprod_[p];
                ratio_of_accepted = largest / magnitude; ties = 0;
            } else {
                ++ties;
                double ratio = largest / magnitude;
                if (ratio < ratio_of_accepted) { chosen = i; ratio_of_accepted = ratio; }
                if (ties >= min_prod * TIES_MULTIPLIER) { pr = pc = chosen; return true; }
            }
        }
        if (!chosen) return false;
        pr = pc = chosen;
        return true;
    }

    bool search_entire_matrix(int step, int& pr, int& pc) {
        bool have = false;
        double largest_mag = 0.0; int lr = 0, lc = 0;
        long min_prod = LARGEST_LONG;
        double ratio_of_accepted = 0.0; long ties = 0;
        for (int col = step; col <= n_; ++col) {
            double largest_in_col = find_largest_in_col_from(col, step);
            if (largest_in_col == 0.0) continue;
            for (int r = step; r <= n_; ++r) {
                if (!ex_[idx(r, col)]) continue;
                double magnitude = mag(v_[idx(r, col)]);
                if (magnitude > largest_mag) { largest_mag = magnitude; lr = r; lc = col; }
                long product = mrow_[r] * mcol_[col];
                if (product <= min_prod && magnitude > rel_threshold_ * largest_in_col &&
                    magnitude > abs_threshold_) {
                    if (product < min_prod) {
                        pr = r; pc = col; have = true; min_prod = product;
                        ratio_of_accepted = largest_in_col / magnitude; ties = 0;
                    } else {
                        ++ties;
                        double ratio = largest_in_col / magnitude;
                        if (ratio < ratio_of_accepted) { pr = r; pc = col; ratio_of_accepted = ratio; }
                        if (ties >= min_prod * TIES_MULTIPLIER) return true;
                    }
                }
            }
        }
        if (have) return true;
        if (largest_mag == 0.0) { error_ = SP_SINGULAR; return false; }
        error_ = SP_SMALL_PIVOT;
        pr = lr; pc = lc;
        return true;
    }

    void swap_rows(int a, int b) {
        if (a == b) return;
        for (int c = 1; c <= n_; ++c) {
            std::swap(v_[idx(a, c)], v_[idx(b, c)]);
            std::swap(ex_[idx(a, c)], ex_[idx(b, c)]);
        }
        std::swap(mrow_[a], mrow_[b]);
        std::swap(i2e_row_[a], i2e_row_[b]);
        e2i_row_[i2e_row_[a]] = a; e2i_row_[i2e_row_[b]] = b;
    }
    void swap_cols(int a, int b) {
        if (a == b) return;
        for (int r = 1; r <= n_; ++r) {
            std::swap(v_[idx(r, a)], v_[idx(r, b)]);
            std::swap(ex_[idx(r, a)], ex_[idx(r, b)]);
        }
        std::swap(mcol_[a], mcol_[b]);
        std::swap(i2e_col_[a], i2e_col_[b]);
        e2i_col_[i2e_col_[a]] = a; e2i_col_[i2e_col_[b]] = b;
    }

    void exchange_rows_and_cols(int row, int col, int step) {
        if (row == step && col == step) return;
        if (row == col) {
            swap_rows(step, row);
            swap_cols(step, col);
            std::swap(mprod_[step], mprod_[row]);
            return;
        }
        long old_step = mprod_[step], old_row = mprod_[row], old_col = mprod_[col];
        if (row != step) {
            swap_rows(step, row);
            mprod_[row] = mrow_[row] * mcol_[row];
            if ((mprod_[row] == 0) != (old_row == 0)) { if (old_row == 0) --singletons_; else ++singletons_; }
        }
        if (col != step) {
            swap_cols(step, col);
            mprod_[col] = mcol_[col] * mrow_[col];
            if ((mprod_[col] == 0) != (old_col == 0)) { if (old_col == 0) --singletons_; else ++singletons_; }
        }
        mprod_[step] = mcol_[step] * mrow_[step];
        if ((mprod_[step] == 0) != (old_step == 0)) { if (old_step == 0) --singletons_; else ++singletons_; }
    }

    void create_fillin(int row, int col) {
        ex_[idx(row, col)] = 1;
        v_[idx(row, col)] = 0.0;
        mprod_[row] = ++mrow_[row] * mcol_[row];
        if (mrow_[row] == 1 && mcol_[row] != 0) --singletons_;
        mprod_[col] = mrow_[col] * ++mcol_[col];
        if (mrow_[col] != 0 && mcol_[col] == 1) --singletons_;
    }

    // RealRowColElimination for the pivot now sitting at (step, step).
    bool real_row_col_elimination(int step) {
        double& piv = v_[idx(step, step)];
        if (std::fabs(piv) == 0.0) { matrix_is_singular(step); return false; }
        piv = 1.0 / piv;
        for (int c = step + 1; c <= n_; ++c) {
            if (!ex_[idx(step, c)]) continue;
            double u = (v_[idx(step, c)] *= piv);
            for (int r = step + 1; r <= n_; ++r) {
                if (!ex_[idx(r, step)]) continue;
                if (!ex_[idx(r, c)]) create_fillin(r, c);
                v_[idx(r, c)] -= u * v_[idx(r, step)];
            }
        }
        return true;
    }

    void update_markowitz_numbers(int step) {
        for (int r = step + 1; r <= n_; ++r) {
            if (!ex_[idx(r, step)]) continue;
            --mrow_[r];
            mprod_[r] = mrow_[r] * mcol_[r];
            if (mrow_[r] == 0) ++singletons_;
        }
        for (int c = step + 1; c <= n_; ++c) {
            if (!ex_[idx(step, c)]) continue;
            --mcol_[c];
            mprod_[c] = mcol_[c] * mrow_[c];
            if (mcol_[c] == 0 && mrow_[c] != 0) ++singletons_;
        }
    }
};

}  // namespace orc
