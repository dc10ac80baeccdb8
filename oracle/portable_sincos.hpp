// oracle/portable_sincos.hpp — TEST INFRASTRUCTURE ONLY.
//
// sin/cos for the CPU oracle.  The reference calls the platform libm (glibc) here — a third-party dependency outside
// /root/reference whose last-bit behaviour is unspecified and decides how the reference's sign-descent falls at flat
// minima.  To make CPU/GPU parity exact, the oracle's default build and the CUDA kernels both
// use one published algorithm: Sun fdlibm 5.3 — __kernel_sin (k_sin.c), __kernel_cos (k_cos.c) and the medium-size
// path of __ieee754_rem_pio2 (e_rem_pio2.c), error < 1 ulp — restated here from its published description.
// Building with -DORACLE_GLIBC_SINCOS (oracle variants "glibc" and "fma") switches back to std::sin / std::cos, i.e. to
// the reference's actual x86-64/glibc behaviour; tests compare the two to show they differ only at ill-conditioned
// points.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {
namespace psc {

inline int32_t hi_word(double x) {
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return (int32_t)(b >> 32);
}
inline double from_hi(int32_t hi) {
    uint64_t b = (uint64_t)(uint32_t)hi << 32;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}

static const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
                    pio2_1t = 6.07710050650619224932e-11, pio2_2 = 6.07710050630396597660e-11,
                    pio2_2t = 2.02226624879595063154e-21, pio2_3 = 2.02226624871116645580e-21,
                    pio2_3t = 8.47842766036889956997e-32;
static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                    S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                    C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

inline double k_sin(double x, double y) {  // k_sin.c, iy = 1
    const double z = x * x;
    const double v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
inline double k_cos(double x, double y) {  // k_cos.c
    const int32_t ix = hi_word(x) & 0x7fffffff;
    const double z = x * x;
    const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    if (ix < 0x3FD33333) return 1.0 - (0.5 * z - (z * r - x * y));
    const double qx = (ix > 0x3fe90000) ? 0.28125 : from_hi(ix - 0x00200000);
    const double hz = 0.5 * z - qx;
    const double a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}

inline void sincos(double x, double &s, double &c) {
#ifdef ORACLE_GLIBC_SINCOS
    s = std::sin(x);
    c = std::cos(x);
#else
    const int32_t hx = hi_word(x);
    const int32_t ix = hx & 0x7fffffff;
    double y0 = x, y1 = 0.0;
    int n = 0;
    if (ix > 0x3fe921fb) {
        if (ix >= 0x412e8480) {  // |x| >= 1e6, inf, nan: outside the medium path
            s = std::sin(x);
            c = std::cos(x);
            return;
        }
        const double t = std::fabs(x);
        n = (int)(t * invpio2 + 0.5);
        const double fn = (double)n;
        double r = t - fn * pio2_1;
        double w = fn * pio2_1t;
        const int j = ix >> 20;
        y0 = r - w;
        int i = j - ((hi_word(y0) >> 20) & 0x7ff);
        if (i > 16) {
            double tt = r;
            w = fn * pio2_2;
            r = tt - w;
            w = fn * pio2_2t - ((tt - r) - w);
            y0 = r - w;
            i = j - ((hi_word(y0) >> 20) & 0x7ff);
            if (i > 49) {
                tt = r;
                w = fn * pio2_3;
                r = tt - w;
                w = fn * pio2_3t - ((tt - r) - w);
                y0 = r - w;
            }
        }
        y1 = (r - y0) - w;
        if (hx < 0) {
            y0 = -y0;
            y1 = -y1;
            n = -n;
        }
    }
    const double ks = k_sin(y0, y1), kc = k_cos(y0, y1);
    switch (n & 3) {
        case 0: s = ks; c = kc; break;
        case 1: s = kc; c = -ks; break;
        case 2: s = -ks; c = -kc; break;
        default: s = -kc; c = ks; break;
    }
#endif
}

}  // namespace psc
}  // namespace oracle
