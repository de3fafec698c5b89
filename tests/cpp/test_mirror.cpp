// Host-side C++ usage of the reference-shaped API (include/tspice_b200.hpp) — the C++ counterpart of the
// reference's cmd/examples/rr/main.go: build the circuit, run OP and transient, read GetResults().
// Compiled (and, on a GPU box, run) by tests/test_cpp_mirror.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include "tspice_b200.hpp"

static const char* RR =
    "* RR Test\n.tran 0.1m 3ms\nVin 1 0 DC 5\nR1 1 2 1k\nR2 2 0 1k";

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    using namespace tsb;
    if (argc > 1 && !std::strcmp(argv[1], "--link-only")) { std::puts(tsb_version()); return 0; }
    try {
        Context ctx(0);
        Circuit ckt = Circuit::FromNetlist(ctx, RR);

        auto op = analysis::NewOP();
        op.Setup(ckt);
        op.Execute();
        auto r = op.GetResults();
        CHECK(r.size() == 3);
        CHECK(r["V(1)"][0] == 5.0 && r["V(2)"][0] == 2.5 && std::fabs(r["I(Vin)"][0] + 2.5e-3) < 1e-18);

        auto tr = analysis::NewTransient(0.0, 3e-3, 1e-4, 1e-4, false);
        tr.Setup(ckt);
        tr.Execute();
        auto w = tr.GetResults();
        CHECK(w.size() == 6 && w["TIME"].size() == 38 && w["TIME"].back() == 0.003);
        for (double v : w["V(2)"]) CHECK(v == 2.5);

        // the batch axis: four values of R2
        Batch batch(ckt, 4);
        batch.SetParam("R2", 0, {500.0, 1000.0, 2000.0, 4000.0});
        auto tb = analysis::NewTransient(0.0, 3e-3, 1e-4, 1e-4, false);
        tb.Setup(batch);
        tb.Execute();
        const double expect[4] = {5.0 * 500 / 1500, 2.5, 5.0 * 2000 / 3000, 4.0};
        for (int i = 0; i < 4; ++i) {
            auto wi = tb.GetResults(i);
            CHECK(std::fabs(wi["V(2)"].back() - expect[i]) < 1e-12);
        }
        for (int32_t st : tb.Status()) CHECK(st == TSB_ST_OK);

        // results on a fixed grid (TSB_OUT_GRID): 300 points at the clamped tStep, constant for this resistive divider
        auto tg = analysis::NewTransient(0.0, 3e-3, 1e-4, 1e-4, false);
        tg.out = TSB_OUT_GRID;
        tg.Setup(batch);
        tg.Execute();
        auto g2 = tg.GetResults(2);
        CHECK(g2["TIME"].size() == 300 && std::fabs(g2["TIME"].front() - 1e-5) < 1e-18 && g2["TIME"].back() == 0.003);
        for (double v : g2["V(2)"]) CHECK(std::fabs(v - expect[2]) < 1e-12);

        // operator level: the rr.cir MNA system itself through the batched LU, strict build
        {
            const std::vector<double> A0 = {1e-3, -1e-3, 1, -1e-3, 2e-3, 0, 1, 0, 0};
            PivotOrder ord = LuOrder(3, A0);
            std::vector<double> A, b;
            for (int q = 0; q < 5; ++q) { A.insert(A.end(), A0.begin(), A0.end()); b.insert(b.end(), {0.0, 0.0, 5.0 + q}); }
            std::vector<int32_t> st;
            auto x = LuSolveBatched(ctx, 3, ord, A, b, st, true);
            for (int q = 0; q < 5; ++q) {
                CHECK(st[q] == 0);
                CHECK(std::fabs(x[3 * q] - (5.0 + q)) < 1e-12 && std::fabs(x[3 * q + 1] - (5.0 + q) / 2) < 1e-12 &&
                      std::fabs(x[3 * q + 2] + (5.0 + q) / 2 * 1e-3) < 1e-15);
            }
        }

        // error behaviour: DC sweep over a source that does not exist (dc.go:64-66), inconsistent lengths (dc.go:21-23)
        bool threw = false;
        try { auto dc = analysis::NewDCSweep({"Vx"}, {0.0}, {1.0}, {0.1}); dc.Setup(ckt); dc.Execute(); } catch (const Error&) { threw = true; }
        CHECK(threw);
        threw = false;
        try { analysis::NewDCSweep({"Vin"}, {0.0, 1.0}, {1.0}, {0.1}); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    std::puts("cpp mirror ok");
    return 0;
}
