// oracle/portable_sincos.hpp — TEST INFRASTRUCTURE ONLY.
//
// sin/cos/atan2 for the CPU oracle.  The reference calls the platform libm (glibc) here — a third-party dependency
// outside /root/reference whose last-bit behaviour is unspecified and decides how the reference's sign-descent falls at
// flat minima.  To make CPU/GPU parity exact, the oracle's default build and the CUDA kernels use
// the same algorithms, made only of IEEE-exact operations (+, -, *, /, fma):
//   sincos : 3-part Cody–Waite reduction of pi/2 with FMA (quadrant by the 1.5*2^52 rounding trick) followed by the
//            published fdlibm 5.3 minimax coefficients (k_sin.c S1..S6, k_cos.c C1..C6) in Horner form with FMA; <= 2 ulp.
//   atan2  : fdlibm 5.3 e_atan2.c / s_atan.c restated from its published description; < 1 ulp.
// Building with -DORACLE_GLIBC_SINCOS (variants "glibc" and "fma") switches back to std::sin / std::cos / std::atan2,
// i.e. to the reference's actual x86-64/glibc behaviour; tests compare the variants to show that they differ only at
// ill-conditioned points.  std::fma must compile to the hardware instruction (-mfma) to be fast; it is exact either way.
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
inline int32_t lo_word(double x) {
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return (int32_t)(uint32_t)b;
}

static const double two_over_pi = 0x1.45f306dc9c883p-1, pio2_hi = 0x1.921fb54442d18p+0, pio2_mid = 0x1.1a62633145c07p-54,
                    pio2_lo = -0x1.f1976b7ed8fbcp-110, magic = 6755399441055744.0;
static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                    S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                    C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

inline void sincos(double x, double &s, double &c) {
#ifdef ORACLE_GLIBC_SINCOS
    s = std::sin(x);
    c = std::cos(x);
#else
    if (!(std::fabs(x) < 1.0e6)) {
        s = std::sin(x);
        c = std::cos(x);
        return;
    }
    const double v = std::fma(x, two_over_pi, magic);
    const int q = lo_word(v);
    const double fn = v - magic;
    double r = std::fma(fn, -pio2_hi, x);
    r = std::fma(fn, -pio2_mid, r);
    r = std::fma(fn, -pio2_lo, r);
    const double z = r * r;
    double ps = std::fma(z, S6, S5);
    ps = std::fma(z, ps, S4);
    ps = std::fma(z, ps, S3);
    ps = std::fma(z, ps, S2);
    ps = std::fma(z, ps, S1);
    const double sr = std::fma(r * z, ps, r);
    double pc = std::fma(z, C6, C5);
    pc = std::fma(z, pc, C4);
    pc = std::fma(z, pc, C3);
    pc = std::fma(z, pc, C2);
    pc = std::fma(z, pc, C1);
    const double cr = std::fma(z * z, pc, std::fma(z, -0.5, 1.0));
    const double ss = (q & 1) ? cr : sr;
    const double cc = (q & 1) ? sr : cr;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
#endif
}

static const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                              -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                              6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                              -3.65315727442169155270e-02, 1.62858201153657823623e-02};
static const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                                 1.57079632679489655800e+00};
static const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                                 6.12323399573676603587e-17};

inline double atan_p(double x) {  // s_atan.c
    const int32_t hx = hi_word(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x44100000) {
        const double zz = atanhi[3] + atanlo[3];
        return (hx > 0) ? zz : -zz;
    }
    if (ix < 0x3fdc0000) {
        if (ix < 0x3e200000) return x;
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3ff30000) {
            if (ix < 0x3fe60000) {
                id = 0;
                x = (2.0 * x - 1.0) / (2.0 + x);
            } else {
                id = 1;
                x = (x - 1.0) / (x + 1.0);
            }
        } else {
            if (ix < 0x40038000) {
                id = 2;
                x = (x - 1.5) / (1.0 + 1.5 * x);
            } else {
                id = 3;
                x = -1.0 / x;
            }
        }
    }
    const double z = x * x;
    const double w = z * z;
    const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const double zz = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (hx < 0) ? -zz : zz;
}

inline double atan2(double y, double x) {  // e_atan2.c
#ifdef ORACLE_GLIBC_SINCOS
    return std::atan2(y, x);
#else
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16, pi_o_2 = 1.5707963267948965580E+00;
    if (!(std::fabs(x) <= 1.79769313486231570815e+308) || !(std::fabs(y) <= 1.79769313486231570815e+308)) return std::atan2(y, x);
    const int32_t hx = hi_word(x), hy = hi_word(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (x == 1.0) return atan_p(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (y == 0.0) {
        switch (m) {
            case 0:
            case 1: return y;
            case 2: return pi;
            default: return -pi;
        }
    }
    if (x == 0.0) return (hy < 0) ? -pi_o_2 : pi_o_2;
    const int k = (iy - ix) >> 20;
    double z;
    if (k > 60) z = pi_o_2 + 0.5 * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0;
    else z = atan_p(std::fabs(y / x));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
#endif
}

}  // namespace psc
}  // namespace oracle
