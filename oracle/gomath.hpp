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

// math.Pow (src/math/pow.go), restated: special cases, Sqrt for y = +-0.5, and the general path
// x**y = Exp(yf*Log(x)) * x**yi with the integer power by repeated squaring of the Frexp mantissa and one Ldexp at
// the end.  For integer y this is a fixed sequence of IEEE multiplications (and one division for y < 0), i.e.
// reproducible bit for bit — unlike libm's pow, which rounds the exact power once.  Exp / Log of the fractional part
// come from libm (<= 1 ulp from Go's).
inline double go_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (std::isnan(x) || std::isnan(y)) return std::nan("");
    if (x == 0) {
        bool odd = std::fabs(y) < 9007199254740992.0 && std::fmod(std::fabs(y), 2.0) == 1.0;
        if (y < 0) return (odd && std::signbit(x)) ? -INFINITY : INFINITY;
        return odd ? x : 0.0;
    }
    if (std::isinf(y)) {
        if (x == -1) return 1;
        if ((std::fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    if (std::isinf(x)) {
        if (x < 0) {
            bool odd = std::fabs(y) < 9007199254740992.0 && std::fmod(std::fabs(y), 2.0) == 1.0;
            if (y < 0) return odd ? -0.0 : 0.0;
            return odd ? -INFINITY : INFINITY;
        }
        return y < 0 ? 0.0 : INFINITY;
    }
    if (y == 0.5) return std::sqrt(x);
    if (y == -0.5) return 1 / std::sqrt(x);
    double yi, yf = std::modf(std::fabs(y), &yi);
    if (yf != 0 && x < 0) return std::nan("");
    if (yi >= 9.223372036854775808e18) {
        if (x == -1) return 1;
        if ((std::fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = std::exp(yf * std::log(x));
    }
    int xe;
    double x1 = std::frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) {
            ae += xe;                                  // overflow / underflow: let Ldexp decide
            break;
        }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    return std::ldexp(a1, ae);
}

// math.Log2 / math.Log10 (src/math/log10.go): log2(x) = Log(frac) * (1/Ln2) + exp with Frexp's split (exact for powers of
// two), log10(x) = log2(x) * (Ln2/Ln10).  Log itself from libm (<= 1 ulp from Go's FreeBSD-derived routine).  They place the
// points of a DEC / OCT frequency sweep (ac.go:100-119).
inline double go_log2(double x) {
    int e;
    double frac = std::frexp(x, &e);
    if (frac == 0.5) return double(e - 1);
    return std::log(frac) * (1 / 0.693147180559945309417232121458176568) + double(e);
}
inline double go_log10(double x) { return go_log2(x) * (0.693147180559945309417232121458176568 / 2.30258509299404568401799145468436421); }

// math.Hypot (src/math/hypot.go; the amd64 assembly follows the same steps): p * Sqrt(1 + (q/p)^2) with p the larger
// magnitude — NOT the correctly rounded hypot of libm.  cmplx.Abs is this (anlysis.go:100).
inline double go_hypot(double p, double q) {
    p = std::fabs(p); q = std::fabs(q);
    if (std::isinf(p) || std::isinf(q)) return INFINITY;
    if (std::isnan(p) || std::isnan(q)) return std::nan("");
    if (p < q) { double t = p; p = q; q = t; }
    if (p == 0) return 0;
    q = q / p;
    return p * std::sqrt(1 + q * q);
}

}  // namespace orc
