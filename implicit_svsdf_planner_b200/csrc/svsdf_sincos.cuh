// svsdf_sincos.cuh — device sin/cos/atan2 used on the SVSDF path.
//
// The reference calls the platform libm here (glibc `sin`/`cos` through Eigen::AngleAxisd, sw_manager.hpp:465-474 and
// back_end_optimizer.hpp:1058-1061; `atan2` in SampleSet2D::initSet :80 and Polygon::isCrossRayOnXDir
// Shape.hpp:1374-1375).  libm is a third-party dependency outside /root/reference whose last-bit behaviour is
// unspecified, and that last bit decides which way the reference's sign-descent falls at flat minima.
// This build therefore pins ONE algorithm, made only of IEEE-exact operations (+, -, *, /, fma), on both sides — the
// CPU oracle carries its own copy in oracle/portable_sincos.hpp — so the strict build reproduces the oracle bit for bit:
//   sincos : 3-part Cody–Waite reduction of pi/2 with FMA (quadrant by the 1.5*2^52 rounding trick), then the fdlibm
//            minimax coefficients (k_sin.c S1..S6, k_cos.c C1..C6) in Horner form with FMA.  <= 2 ulp on |x| < 1e6,
//            22 FP64 instructions (the libm-style table/branch code paths are avoided on purpose).
//   atan2  : fdlibm e_atan2.c / s_atan.c (argument reduction to 4 intervals, 11-term polynomial), < 1 ulp.
// |x| >= 1e6 or non-finite arguments fall back to the CUDA library (never reached by yaw angles).
// Coefficients sit in __constant__ memory so FP64 instructions take them as constant-bank operands.
#pragma once
#include <cuda_runtime.h>

namespace svsdf {
namespace dev {

static __constant__ double SC_TAB[20] = {
    /* 0 2/pi   */ 0x1.45f306dc9c883p-1,
    /* 1 pio2_hi*/ 0x1.921fb54442d18p+0, /* 2 pio2_mid */ 0x1.1a62633145c07p-54, /* 3 pio2_lo */ -0x1.f1976b7ed8fbcp-110,
    /* 4 magic  */ 6755399441055744.0,   /* 1.5 * 2^52 */
    /* 5 S1 */ -1.66666666666666324348e-01, /* 6 S2 */ 8.33333333332248946124e-03, /* 7 S3 */ -1.98412698298579493134e-04,
    /* 8 S4 */ 2.75573137070700676789e-06, /* 9 S5 */ -2.50507602534068634195e-08, /* 10 S6 */ 1.58969099521155010221e-10,
    /* 11 C1 */ 4.16666666666666019037e-02, /* 12 C2 */ -1.38888888888741095749e-03, /* 13 C3 */ 2.48015872894767294178e-05,
    /* 14 C4 */ -2.75573143513906633035e-07, /* 15 C5 */ 2.08757232129817482790e-09, /* 16 C6 */ -1.13596475577881948265e-11,
    /* 17 */ -0.5, /* 18 */ 1.0, /* 19 */ 0.0};

__device__ __forceinline__ void sincos_portable(double x, double &s, double &c) {
    if (!(fabs(x) < 1.0e6)) {  // also catches NaN / inf
        ::sincos(x, &s, &c);
        return;
    }
    const double v = fma(x, SC_TAB[0], SC_TAB[4]);  // x * 2/pi + 1.5*2^52: integer part lands in the low mantissa bits
    const int q = __double2loint(v);
    const double fn = v - SC_TAB[4];
    double r = fma(fn, -SC_TAB[1], x);
    r = fma(fn, -SC_TAB[2], r);
    r = fma(fn, -SC_TAB[3], r);
    const double z = r * r;
    double ps = fma(z, SC_TAB[10], SC_TAB[9]);
    ps = fma(z, ps, SC_TAB[8]);
    ps = fma(z, ps, SC_TAB[7]);
    ps = fma(z, ps, SC_TAB[6]);
    ps = fma(z, ps, SC_TAB[5]);
    const double sr = fma(r * z, ps, r);
    double pc = fma(z, SC_TAB[16], SC_TAB[15]);
    pc = fma(z, pc, SC_TAB[14]);
    pc = fma(z, pc, SC_TAB[13]);
    pc = fma(z, pc, SC_TAB[12]);
    pc = fma(z, pc, SC_TAB[11]);
    const double cr = fma(z * z, pc, fma(z, SC_TAB[17], SC_TAB[18]));
    const double ss = (q & 1) ? cr : sr;
    const double cc = (q & 1) ? sr : cr;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// fdlibm s_atan.c
static __constant__ double AT_TAB[20] = {
    /* aT[0..10] */ 3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01,
    -1.11111104054623557880e-01, 9.09088713343650656196e-02, -7.69187620504482999495e-02, 6.66107313738753120669e-02,
    -5.83357013379057348645e-02, 4.97687799461593236017e-02, -3.65315727442169155270e-02, 1.62858201153657823623e-02,
    /* 11..14 atanhi */ 4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
    1.57079632679489655800e+00,
    /* 15..18 atanlo */ 2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
    6.12323399573676603587e-17, 0.0};

__device__ __forceinline__ double atan_portable(double x) {
    const int hx = __double2hiint(x);
    const int ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x44100000) {  // |x| >= 2^66 (or inf; NaN handled by the caller)
        const double zz = AT_TAB[14] + AT_TAB[18];
        return (hx > 0) ? zz : -zz;
    }
    if (ix < 0x3fdc0000) {  // |x| < 0.4375
        if (ix < 0x3e200000) return x;  // |x| < 2^-29
        id = -1;
    } else {
        x = fabs(x);
        if (ix < 0x3ff30000) {      // |x| < 1.1875
            if (ix < 0x3fe60000) {  // 7/16 <= |x| < 11/16
                id = 0;
                x = (2.0 * x - 1.0) / (2.0 + x);
            } else {  // 11/16 <= |x| < 19/16
                id = 1;
                x = (x - 1.0) / (x + 1.0);
            }
        } else {
            if (ix < 0x40038000) {  // |x| < 2.4375
                id = 2;
                x = (x - 1.5) / (1.0 + 1.5 * x);
            } else {  // 2.4375 <= |x| < 2^66
                id = 3;
                x = -1.0 / x;
            }
        }
    }
    const double z = x * x;
    const double w = z * z;
    const double s1 = z * (AT_TAB[0] + w * (AT_TAB[2] + w * (AT_TAB[4] + w * (AT_TAB[6] + w * (AT_TAB[8] + w * AT_TAB[10])))));
    const double s2 = w * (AT_TAB[1] + w * (AT_TAB[3] + w * (AT_TAB[5] + w * (AT_TAB[7] + w * AT_TAB[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const double zz = AT_TAB[11 + id] - ((x * (s1 + s2) - AT_TAB[15 + id]) - x);
    return (hx < 0) ? -zz : zz;
}

// fdlibm e_atan2.c (finite arguments; anything else goes to the CUDA library)
__device__ __forceinline__ double atan2_portable(double y, double x) {
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16, pi_o_2 = 1.5707963267948965580E+00;
    if (!(fabs(x) <= 1.79769313486231570815e+308) || !(fabs(y) <= 1.79769313486231570815e+308)) return ::atan2(y, x);
    const int hx = __double2hiint(x), hy = __double2hiint(y);
    const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (x == 1.0) return atan_portable(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);  // 2*sign(x) + sign(y)
    if (y == 0.0) {
        switch (m) {
            case 0:
            case 1: return y;  // atan(+-0, +anything) = +-0
            case 2: return pi;
            default: return -pi;
        }
    }
    if (x == 0.0) return (hy < 0) ? -pi_o_2 : pi_o_2;
    const int k = (iy - ix) >> 20;
    double z;
    if (k > 60) z = pi_o_2 + 0.5 * pi_lo;            // |y/x| > 2^60
    else if (hx < 0 && k < -60) z = 0.0;             // |y|/x < -2^60
    else z = atan_portable(fabs(y / x));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

}  // namespace dev
}  // namespace svsdf
