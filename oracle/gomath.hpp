// ORACLE — TEST INFRASTRUCTURE ONLY (see sparse13.hpp header).
//
// gomath.hpp — the math-library behaviour the reference's device models see.
// The reference is Go; its math.Sin is the pure-Go Cephes routine on amd64, while the CPU
// oracle is C++.  Sin feeds every SIN source (pkg/device/vsource.go:117-119,
// isource.go:115-117), so it is restated here from the published Go standard library
// algorithm (src/math/sin.go: Cody-Waite reduction with PI4A/B/C, degree-6 minimax
// polynomials in z^2) to keep linear transient runs as close to a Go run as possible.
// math.Exp / math.Log / math.Pow have architecture-specific Go implementations that are not
// restated; libm is used (<= 1 ulp difference, far inside the 1e-9 parity tolerance).
// math.Mod is exact in every correct implementation (fmod).
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

inline double go_sin(double x) {
    static const double sc[6] = {
        1.58962301576546568060e-10, -2.50507477628578072866e-8, 2.75573136213857245213e-6,
        -1.98412698295895385996e-4, 8.33333333332211858878e-3, -1.66666666666666307295e-1};
    static const double cc[6] = {
        -1.13585365213876817300e-11, 2.08757008419747316778e-9, -2.75573141792967388112e-7,
        2.48015872888517045348e-5, -1.38888888888730564116e-3, 4.16666666666665929218e-2};
    const double PI4A = 7.85398125648498535156e-1;
    const double PI4B = 3.77489470793079817668e-8;
    const double PI4C = 2.69515142907905952645e-15;
    const double M4PI = 1.2732395447351626861510701069801148;  // 4/Pi (Go: exact constant, rounded once)
    if (x == 0.0 || std::isnan(x)) return x;
    if (std::isinf(x)) return std::nan("");
    bool sign = false;
    if (x < 0) { x = -x; sign = true; }
    if (x >= 536870912.0) {           // reduceThreshold = 1<<29: Payne-Hanek in Go; not reached here
        double y = std::sin(x);
        return sign ? -y : y;
    }
    uint64_t j = (uint64_t)(x * M4PI);
    double y = (double)j;
    if (j & 1) { j++; y++; }
    j &= 7;
    double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
    if (j > 3) { sign = !sign; j -= 4; }
    double zz = z * z;
    if (j == 1 || j == 2) {
        y = 1.0 - 0.5 * zz + zz * zz * ((((((cc[0] * zz) + cc[1]) * zz + cc[2]) * zz + cc[3]) * zz + cc[4]) * zz + cc[5]);
    } else {
        y = z + z * zz * ((((((sc[0] * zz) + sc[1]) * zz + sc[2]) * zz + sc[3]) * zz + sc[4]) * zz + sc[5]);
    }
    return sign ? -y : y;
}

// math.Max / math.Min (src/math/dim.go): +Inf / -Inf win over NaN, otherwise NaN propagates (C's fmax / fmin DROP
// a NaN operand), and Max(+0, -0) = +0, Min(+0, -0) = -0.
inline double go_max(double x, double y) {
    if ((std::isinf(x) && x > 0) || (std::isinf(y) && y > 0)) return INFINITY;
    if (std::isnan(x) || std::isnan(y)) return std::nan("");
    if (x == 0 && x == y) return std::signbit(x) ? y : x;
    return x > y ? x : y;
}
inline double go_min(double x, double y) {
    if ((std::isinf(x) && x < 0) || (std::isinf(y) && y < 0)) return -INFINITY;
    if (std::isnan(x) || std::isnan(y)) return std::nan("");
    if (x == 0 && x == y) return std::signbit(x) ? x : y;
    return x < y ? x : y;
}

}  // namespace orc
