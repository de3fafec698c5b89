// hostlu.hpp — host symbolic pass: the pivot order of the reference's first factorisation.
//
// The north star fixes the GPU pivot order to "the reference's symbolic pass".  The reference
// gets it from github.com/edp1096/sparse (a Go port of Sparse 1.3, not vendored; call sites
// pkg/matrix/circuit.go:20-33,126-150): the first Factor() orders-and-factors with Markowitz
// counts, a relative threshold of 1e-3 and diagonal preference, and every later Factor()
// reuses that order.  This class replays that first factorisation ONCE, on the nominal
// instance, on the host.  It is not a solver path: batches are only ever solved on the GPU.
//
// Internal numbering follows Sparse's Translate (first touch, row before column) because the
// diagonal search scans internal indices from high to low with Diag[Step] inspected first.
#pragma once
#include <cmath>
#include <vector>

namespace tsb {

class MarkowitzLU {
public:
    explicit MarkowitzLU(int n) : n_(n), a_((n + 1) * (n + 1)), on_((n + 1) * (n + 1), 0),
        e2i_(n + 1, -1), rext_(n + 1), cext_(n + 1), mr_(n + 2), mc_(n + 2), mp_(n + 3) {
        for (int i = 0; i <= n; ++i) { rext_[i] = i; cext_[i] = i; }
        e2i_[0] = 0;
    }
    // AddElement: creates the element at first touch (Translate: row before column), accumulates.
    void create(int er, int ec) { int i = intern(er); int j = intern(ec); on_[at(i, j)] = 1; }
    void add(int er, int ec, double v) { create(er, ec); a_[at(e2i_[er], e2i_[ec])] += v; }
    int ext2int(int e) const { return e2i_[e]; }
    int assigned() const { return next_; }

    // First Factor(): order and eliminate.  Returns false if the matrix is singular.
    bool order_and_factor() {
        for (int i = 1; i <= n_; ++i) {
            long cr = -1, cc = -1;
            for (int j = 1; j <= n_; ++j) { if (on_[at(i, j)]) ++cr; if (on_[at(j, i)]) ++cc; }
            mr_[i] = cr; mc_[i] = cc;
        }
        singletons_ = 0;
        for (int i = 1; i <= n_; ++i) if ((mp_[i] = mr_[i] * mc_[i]) == 0) ++singletons_;
        for (int k = 1; k <= n_; ++k) {
            int r = 0, c = 0;
            if (!pick(k, r, c)) return false;
            bring_to(k, r, c);
            double& piv = a_[at(k, k)];
            if (std::fabs(piv) == 0.0) return false;
            piv = 1.0 / piv;
            for (int j = k + 1; j <= n_; ++j) {
                if (!on_[at(k, j)]) continue;
                double u = (a_[at(k, j)] *= piv);
                for (int i = k + 1; i <= n_; ++i) {
                    if (!on_[at(i, k)]) continue;
                    if (!on_[at(i, j)]) fill(i, j);
                    a_[at(i, j)] -= u * a_[at(i, k)];
                }
            }
            for (int i = k + 1; i <= n_; ++i) if (on_[at(i, k)]) { mp_[i] = --mr_[i] * mc_[i]; if (mr_[i] == 0) ++singletons_; }
            for (int j = k + 1; j <= n_; ++j) if (on_[at(k, j)]) { --mc_[j]; mp_[j] = mc_[j] * mr_[j]; if (mc_[j] == 0 && mr_[j] != 0) ++singletons_; }
        }
        return true;
    }
    // Solve with the factors (1-based external vectors, x[0] = 0).
    void solve(const std::vector<double>& b, std::vector<double>& x) const {
        std::vector<double> c(n_ + 1, 0.0);
        for (int i = 1; i <= n_; ++i) c[i] = b[rext_[i]];
        for (int i = 1; i <= n_; ++i) {
            double t = c[i];
            if (t == 0.0) continue;
            c[i] = (t *= a_[at(i, i)]);
            for (int r = i + 1; r <= n_; ++r) if (on_[at(r, i)]) c[r] -= t * a_[at(r, i)];
        }
        for (int i = n_; i >= 1; --i) {
            double t = c[i];
            for (int j = i + 1; j <= n_; ++j) if (on_[at(i, j)]) t -= a_[at(i, j)] * c[j];
            c[i] = t;
        }
        x.assign(n_ + 1, 0.0);
        for (int i = 1; i <= n_; ++i) x[cext_[i]] = c[i];
    }
    int pivot_row(int k) const { return rext_[k]; }
    int pivot_col(int k) const { return cext_[k]; }

private:
    int n_, next_ = 0;
    std::vector<double> a_;
    std::vector<char> on_;
    std::vector<int> e2i_, rext_, cext_;     // ext -> internal (creation), internal -> ext row / col
    std::vector<long> mr_, mc_, mp_;
    int singletons_ = 0;
    static constexpr double kRel = 1e-3, kAbs = 0.0;
    static constexpr long kTies = 5, kBig = 0x7fffffffL;

    int at(int r, int c) const { return r * (n_ + 1) + c; }
    int fresh(int e) { e2i_[e] = ++next_; rext_[next_] = e; cext_[next_] = e; return next_; }
    int intern(int er) { int i = e2i_[er]; return i == -1 ? fresh(er) : i; }
    void fill(int i, int j) {
        on_[at(i, j)] = 1; a_[at(i, j)] = 0.0;
        mp_[i] = ++mr_[i] * mc_[i];
        if (mr_[i] == 1 && mc_[i] != 0) --singletons_;
        mp_[j] = mr_[j] * ++mc_[j];
        if (mr_[j] != 0 && mc_[j] == 1) --singletons_;
    }
    double colmax_excluding(int row, int col, int k) const {
        double big = 0.0;
        for (int r = k; r <= n_; ++r) if (on_[at(r, col)] && r != row) big = std::fmax(big, std::fabs(a_[at(r, col)]));
        return big;
    }
    bool ok_pivot(int r, int c, int k) const {
        double m = std::fabs(a_[at(r, c)]);
        return m > kAbs && m > kRel * colmax_excluding(r, c, k);
    }
    bool pick(int k, int& r, int& c) {
        if (singletons_ && pick_singleton(k, r, c)) return true;
        if (pick_diag_quick(k, r, c)) return true;
        if (pick_diag_careful(k, r, c)) return true;
        return pick_anywhere(k, r, c);
    }
    bool pick_singleton(int k, int& pr, int& pc) {
        mp_[n_ + 1] = mp_[k];
        int left = singletons_--;
        mp_[k - 1] = 0;
        int p = n_ + 1;
        while (left-- > 0) {
            while (mp_[p--] != 0) {}
            int i = p + 1;
            if (i < k) break;
            if (i > n_) i = k;
            if (on_[at(i, i)]) {
                if (ok_pivot(i, i, k)) { pr = pc = i; return true; }
            } else if (mc_[i] == 0) {
                int r = k; while (r <= n_ && !on_[at(r, i)]) ++r;
                if (r > n_) break;
                if (ok_pivot(r, i, k)) { pr = r; pc = i; return true; }
                if (mr_[i] == 0) {
                    int c = k; while (c <= n_ && !on_[at(i, c)]) ++c;
                    if (c > n_) break;
                    if (ok_pivot(i, c, k)) { pr = i; pc = c; return true; }
                }
            } else {
                int c = k; while (c <= n_ && !on_[at(i, c)]) ++c;
                if (c > n_) break;
                if (ok_pivot(i, c, k)) { pr = i; pc = c; return true; }
            }
        }
        ++singletons_;
        return false;
    }
    bool pick_diag_quick(int k, int& pr, int& pc) {
        int best = 0; long lo = kBig;
        mp_[n_ + 1] = mp_[k];
        mp_[k - 1] = -1;
        int p = n_ + 2;
        for (;;) {
            while (mp_[--p] >= lo) {}
            int i = p;
            if (i < k) break;
            if (i > n_) i = k;
            if (!on_[at(i, i)]) continue;
            double m = std::fabs(a_[at(i, i)]);
            if (m <= kAbs) continue;
            if (mp_[p] == 1) {
                int oc = 0, orow = 0;
                for (int j = i + 1; j <= n_ && !oc; ++j) if (on_[at(i, j)]) oc = j;
                for (int r = i + 1; r <= n_ && !orow; ++r) if (on_[at(r, i)]) orow = r;
                if (!oc && !orow) {
                    for (int j = k; j <= n_ && !oc; ++j) if (j != i && on_[at(i, j)]) oc = j;
                    for (int r = k; r <= n_ && !orow; ++r) if (r != i && on_[at(r, i)]) orow = r;
                }
                if (oc && orow && oc == orow && m >= std::fmax(std::fabs(a_[at(i, oc)]), std::fabs(a_[at(orow, i)]))) {
                    pr = pc = i; return true;
                }
            }
            lo = mp_[p];
            best = i;
        }
        if (best && std::fabs(a_[at(best, best)]) <= kRel * colmax_excluding(best, best, k)) best = 0;
        if (!best) return false;
        pr = pc = best;
        return true;
    }
    bool pick_diag_careful(int k, int& pr, int& pc) {
        int best = 0; long lo = kBig, ties = 0; double best_ratio = 0;
        mp_[n_ + 1] = mp_[k];
        for (int j = n_ + 1; j > k; --j) {
            if (mp_[j] > lo) continue;
            int i = j > n_ ? k : j;
            if (!on_[at(i, i)]) continue;
            double m = std::fabs(a_[at(i, i)]);
            if (m <= kAbs) continue;
            double big = colmax_excluding(i, i, k);
            if (m <= kRel * big) continue;
            if (mp_[j] < lo) { best = i; lo = mp_[j]; best_ratio = big / m; ties = 0; }
            else {
                ++ties;
                double ratio = big / m;
                if (ratio < best_ratio) { best = i; best_ratio = ratio; }
                if (ties >= lo * kTies) break;
            }
        }
        if (!best) return false;
        pr = pc = best;
        return true;
    }
    bool pick_anywhere(int k, int& pr, int& pc) {
        bool have = false; long lo = kBig, ties = 0; double best_ratio = 0, biggest = 0; int br = 0, bc = 0;
        for (int col = k; col <= n_; ++col) {
            double colbig = 0;
            for (int r = k; r <= n_; ++r) if (on_[at(r, col)]) colbig = std::fmax(colbig, std::fabs(a_[at(r, col)]));
            if (colbig == 0.0) continue;
            for (int r = k; r <= n_; ++r) {
                if (!on_[at(r, col)]) continue;
                double m = std::fabs(a_[at(r, col)]);
                if (m > biggest) { biggest = m; br = r; bc = col; }
                long prod = mr_[r] * mc_[col];
                if (prod <= lo && m > kRel * colbig && m > kAbs) {
                    if (prod < lo) { pr = r; pc = col; have = true; lo = prod; best_ratio = colbig / m; ties = 0; }
                    else {
                        ++ties;
                        double ratio = colbig / m;
                        if (ratio < best_ratio) { pr = r; pc = col; best_ratio = ratio; }
                        if (ties >= lo * kTies) return true;
                    }
                }
            }
        }
        if (have) return true;
        if (biggest == 0.0) return false;
        pr = br; pc = bc;
        return true;
    }
    void swap_rows(int x, int y) {
        if (x == y) return;
        for (int j = 1; j <= n_; ++j) { std::swap(a_[at(x, j)], a_[at(y, j)]); std::swap(on_[at(x, j)], on_[at(y, j)]); }
        std::swap(mr_[x], mr_[y]); std::swap(rext_[x], rext_[y]);
    }
    void swap_cols(int x, int y) {
        if (x == y) return;
        for (int i = 1; i <= n_; ++i) { std::swap(a_[at(i, x)], a_[at(i, y)]); std::swap(on_[at(i, x)], on_[at(i, y)]); }
        std::swap(mc_[x], mc_[y]); std::swap(cext_[x], cext_[y]);
    }
    void fix_singletons(long before, long after) { if ((after == 0) != (before == 0)) singletons_ += before == 0 ? -1 : 1; }
    void bring_to(int k, int r, int c) {
        if (r == k && c == k) return;
        if (r == c) { swap_rows(k, r); swap_cols(k, c); std::swap(mp_[k], mp_[r]); return; }
        long ok = mp_[k], orr = mp_[r], oc = mp_[c];
        if (r != k) { swap_rows(k, r); mp_[r] = mr_[r] * mc_[r]; fix_singletons(orr, mp_[r]); }
        if (c != k) { swap_cols(k, c); mp_[c] = mc_[c] * mr_[c]; fix_singletons(oc, mp_[c]); }
        mp_[k] = mc_[k] * mr_[k]; fix_singletons(ok, mp_[k]);
    }
};

}  // namespace tsb
