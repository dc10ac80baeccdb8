// svsdf_sincos.cuh — device sin/cos used on the SVSDF path.
//
// The reference calls the platform libm (glibc `sin`/`cos` through Eigen::AngleAxisd, sw_manager.hpp:465-474, and
// back_end_optimizer.hpp:1058-1061).  libm is a third-party dependency outside /root/reference, and its last-bit
// behaviour decides which way the reference's sign-descent falls at flat minima, so this build
// pins one published algorithm on both sides: Sun fdlibm 5.3 (k_sin.c, k_cos.c and the medium-size argument path of
// e_rem_pio2.c; error < 1 ulp).  The CPU oracle carries its own copy (oracle/portable_sincos.hpp); with
// -fmad=false the two produce identical bits, which is what makes the strict parity tests exact.
// Only +, -, * (no division, no table look-ups): 3-stage Cody–Waite reduction with the 33+33+53-bit split of pi/2,
// then the degree-13 / degree-14 minimax kernels.  |x| >= 1e6 (never reached by a yaw angle) falls back to CUDA's
// sincos.  Coefficients are read from __constant__ memory so FP64 instructions take them as constant-bank operands.
#pragma once
#include <cuda_runtime.h>

namespace svsdf {
namespace dev {

static __constant__ double SC_TAB[20] = {
    /* 0 invpio2 */ 6.36619772367581382433e-01,  /* 1 pio2_1  */ 1.57079632673412561417e+00,
    /* 2 pio2_1t */ 6.07710050650619224932e-11,  /* 3 pio2_2  */ 6.07710050630396597660e-11,
    /* 4 pio2_2t */ 2.02226624879595063154e-21,  /* 5 pio2_3  */ 2.02226624871116645580e-21,
    /* 6 pio2_3t */ 8.47842766036889956997e-32,
    /* 7 S1 */ -1.66666666666666324348e-01, /* 8 S2 */ 8.33333333332248946124e-03, /* 9 S3 */ -1.98412698298579493134e-04,
    /* 10 S4 */ 2.75573137070700676789e-06, /* 11 S5 */ -2.50507602534068634195e-08, /* 12 S6 */ 1.58969099521155010221e-10,
    /* 13 C1 */ 4.16666666666666019037e-02, /* 14 C2 */ -1.38888888888741095749e-03, /* 15 C3 */ 2.48015872894767294178e-05,
    /* 16 C4 */ -2.75573143513906633035e-07, /* 17 C5 */ 2.08757232129817482790e-09, /* 18 C6 */ -1.13596475577881948265e-11,
    /* 19 */ 0.0};

// fdlibm __kernel_sin(x, y, iy = 1) and __kernel_cos(x, y) on |x| <~ pi/4 with tail y
__device__ __forceinline__ double k_sin(double x, double y) {
    const double z = x * x;
    const double v = z * x;
    const double r = SC_TAB[8] + z * (SC_TAB[9] + z * (SC_TAB[10] + z * (SC_TAB[11] + z * SC_TAB[12])));
    return x - ((z * (0.5 * y - v * r) - y) - v * SC_TAB[7]);
}
__device__ __forceinline__ double k_cos(double x, double y) {
    const int ix = __double2hiint(x) & 0x7fffffff;
    const double z = x * x;
    const double r = z * (SC_TAB[13] + z * (SC_TAB[14] + z * (SC_TAB[15] + z * (SC_TAB[16] + z * (SC_TAB[17] + z * SC_TAB[18])))));
    if (ix < 0x3FD33333) return 1.0 - (0.5 * z - (z * r - x * y));
    const double qx = (ix > 0x3fe90000) ? 0.28125 : __hiloint2double(ix - 0x00200000, 0);
    const double hz = 0.5 * z - qx;
    const double a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}

__device__ __forceinline__ void sincos_fdlibm(double x, double &s, double &c) {
    const int hx = __double2hiint(x);
    const int ix = hx & 0x7fffffff;
    double y0 = x, y1 = 0.0;
    int n = 0;
    if (ix > 0x3fe921fb) {  // |x| > ~pi/4
        if (ix >= 0x412e8480) {  // |x| >= 1e6 (or inf/nan): outside the Cody–Waite range
            ::sincos(x, &s, &c);
            return;
        }
        const double t = fabs(x);
        n = (int)(t * SC_TAB[0] + 0.5);
        const double fn = (double)n;
        double r = t - fn * SC_TAB[1];
        double w = fn * SC_TAB[2];  // 1st round good to 85 bit
        const int j = ix >> 20;
        y0 = r - w;
        int i = j - ((__double2hiint(y0) >> 20) & 0x7ff);
        if (i > 16) {  // 2nd iteration needed, good to 118
            double tt = r;
            w = fn * SC_TAB[3];
            r = tt - w;
            w = fn * SC_TAB[4] - ((tt - r) - w);
            y0 = r - w;
            i = j - ((__double2hiint(y0) >> 20) & 0x7ff);
            if (i > 49) {  // 3rd iteration, 151 bits
                tt = r;
                w = fn * SC_TAB[5];
                r = tt - w;
                w = fn * SC_TAB[6] - ((tt - r) - w);
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
}

}  // namespace dev
}  // namespace svsdf
