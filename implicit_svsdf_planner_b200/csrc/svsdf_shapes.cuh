// svsdf_shapes.cuh — device-side robot-shape SDF functors (2-D, FP64), one template specialisation per
// registry key of the reference (src/utils/include/utils/Shape.hpp; line of each value function cited).
// Operation order follows the reference so that a -fmad=false build differs from the CPU only through
// libm (sin/cos); every function takes the body-frame point AFTER the shape pre-transform.
#pragma once
#include "svsdf_sincos.cuh"
#include "svsdf_types.h"

namespace svsdf {
namespace dev {

// std::max / std::min exactly as libstdc++ defines them (the reference calls these, not fmax/fmin): 1 compare + select
__device__ __forceinline__ double smaxd(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double smind(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double clipd(double v, double lo, double hi) { return smaxd(smind(v, hi), lo); }
__device__ __forceinline__ double len2(double x, double y) { return sqrt(x * x + y * y); }

// a / b for a divisor b known at compile time, given rb = RN(1/b): q = RN(a*rb), r = a - b*q (exact, FMA),
// result RN(q + r*rb).  Markstein's correction step: returns the correctly rounded quotient (bit-identical to the
// IEEE division the reference performs; tests/test_gpu_parity.py::test_strict_shape_functors_are_bitwise checks it)
// in 3 FP64 instructions instead of the ~20-instruction generic division sequence.
__device__ __forceinline__ double div_const(double a, double b, double rb) {
    const double q = a * rb;
    const double r = fma(-b, q, a);
    return fma(r, rb, q);
}

// Literal constants of the shape functors live in __constant__ memory so that FP64 instructions read them as
// constant-bank operands (a double immediate otherwise costs two UMOV per use).  Values are the reference's literals;
// derived entries are folded by the host compiler in IEEE double, one rounding per operation, exactly as the
// reference computes them at run time.
namespace kc {
constexpr double star_k1x = 0.809016994375, star_k1y = -0.587785252292, star_r = 2.8, star_rf = 0.6;
constexpr double star_bax = star_rf * (-star_k1y) - 0.0, star_bay = star_rf * star_k1x - 1.0;
constexpr double star_bb = star_bax * star_bax + star_bay * star_bay;
constexpr double trap_k2x = 3.0 - 1.0, trap_k2y = 2.0 * 2.0;
constexpr double trap_kk = trap_k2x * trap_k2x + trap_k2y * trap_k2y;
constexpr double rhom_bb = 1.0 * 1.0 + 4.5 * 4.5;
}  // namespace kc
__constant__ double KSTAR[8] = {kc::star_k1x, kc::star_k1y, kc::star_r, kc::star_bax, kc::star_bay, kc::star_bb,
                                1.0 / kc::star_bb, 2.0};
__constant__ double KHORSE[4] = {1.5, 1.55, 0.20, 0.0};
__constant__ double KMISC[16] = {
    /*0 pie r*/ 3.0, /*1 arc ra*/ 2.3333, /*2 arc rb*/ 0.5, /*3 tunnel whx*/ 2.5, /*4 tunnel why*/ 1.5,
    /*5 cutdisk r*/ 5.0, /*6 cutdisk h*/ 2.0, /*7 trap kk*/ kc::trap_kk, /*8 1/trap kk*/ 1.0 / kc::trap_kk,
    /*9 rhombus by*/ 4.5, /*10 rhombus bb*/ kc::rhom_bb, /*11 1/bb*/ 1.0 / kc::rhom_bb, /*12*/ 0.25, /*13*/ 0.75,
    /*14*/ 0.5, /*15*/ 2.4};

template <int SHAPE>
struct ShapeFn;

// star — Shape.hpp:584-601 (r = 2.8, rf = 0.6)
template <>
struct ShapeFn<SH_STAR> {
    static __device__ __forceinline__ double sdf(const ShapeParams &, double px, double py) {
        const double k1x = KSTAR[0], k1y = KSTAR[1], r = KSTAR[2];
        const double k2x = -k1x, k2y = k1y;
        px = fabs(px);
        double m = KSTAR[7] * smaxd(k1x * px + k1y * py, 0.0);
        px -= m * k1x;
        py -= m * k1y;
        m = KSTAR[7] * smaxd(k2x * px + k2y * py, 0.0);
        px -= m * k2x;
        py -= m * k2y;
        px = fabs(px);
        py -= r;
        const double bax = KSTAR[3], bay = KSTAR[4];  // rf * (-k1.y, k1.x) - (0, 1)
        double h = clipd(div_const(px * bax + py * bay, KSTAR[5], KSTAR[6]), 0.0, r);
        double dx = px - bax * h, dy = py - bay * h;
        return len2(dx, dy) * copysign(1.0, py * bax - px * bay);
    }
};

// sdHorseshoe — Shape.hpp:870-891; cst = (cos 20.5, sin 20.5) computed on the host (:855)
template <>
struct ShapeFn<SH_HORSESHOE> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        const double r = KHORSE[0], wx = KHORSE[1], wy = KHORSE[2];
        const double cx = S.cst[0], cy = S.cst[1];
        px = fabs(px);
        double l = len2(px, py);
        double qx = -cx * px + cy * py;
        double qy = cy * px + cx * py;
        double px0 = qx;
        if (px0 <= 0 && qy <= 0) qx = l * copysign(1.0, -cx);
        if (px0 <= 0) qy = l;
        qx = qx - wx;
        qy = fabs(qy - r) - wy;
        double tx = smaxd(qx, 0.0), ty = smaxd(qy, 0.0);
        return len2(tx, ty) + smind(0.0, smaxd(qx, qy));
    }
};

// sdPie / sdPie2 — Shape.hpp:1253-1260 / 1294-1301; cst = (cos 43, sin 43) / (cos 1, sin 1)
__device__ __forceinline__ double sd_pie_c(double px, double py, double cx, double cy) {
    const double r = 3.0;
    px = fabs(px);
    double l = len2(px, py) - r;
    double k = clipd(px * cx + py * cy, 0.0, r);
    double dx = px - cx * k, dy = py - cy * k;
    double m = len2(dx, dy);
    return smaxd(l, m * copysign(1.0, cy * px - cx * py));
}
template <>
struct ShapeFn<SH_PIE> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        return sd_pie_c(px, py, S.cst[0], S.cst[1]);
    }
};
template <>
struct ShapeFn<SH_PIE2> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        return sd_pie_c(px, py, S.cst[0], S.cst[1]);
    }
};

// sdArc — Shape.hpp:1334-1343; cst = (sin 20, cos 20), ra = 2.3333, rb = 0.5
template <>
struct ShapeFn<SH_ARC> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        const double scx = S.cst[0], scy = S.cst[1];
        const double ra = 2.3333, rb = 0.5;
        px = fabs(px);
        bool cond = scy * px > scx * py;
        double ax = px - scx * ra, ay = py - scy * ra;
        double dist1 = len2(ax, ay);
        double dist2 = fabs(len2(px, py) - ra);
        return (cond ? dist1 : dist2) - rb;
    }
};

// sdTunnel — Shape.hpp:642-658; wh = (2.5, 1.5)
template <>
struct ShapeFn<SH_TUNNEL> {
    static __device__ __forceinline__ double sdf(const ShapeParams &, double px, double py) {
        const double whx = 2.5, why = 1.5;
        px = fabs(px);
        py = -py;
        double qx = px - whx, qy = py - why;
        double mq = smaxd(qx, 0.0);
        double d1 = mq * mq + qy * qy;
        qx = (py > 0.0) ? qx : len2(px, py) - whx;
        double mqy = smaxd(qy, 0.0);
        double d2 = qx * qx + mqy * mqy;
        double d = sqrt(smind(d1, d2));
        return (smaxd(qx, qy) < 0.0) ? -d : d;
    }
};

// sdCutDisk — Shape.hpp:698-711; r = 5, h = 2; cst[0] = sqrt(r*r - h*h)
template <>
struct ShapeFn<SH_CUTDISK> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        const double r = 5.0, h = 2.0;
        const double w = S.cst[0];
        px = fabs(px);
        double s = smaxd((h - r) * px * px + w * w * (h + r - 2.0 * py), h * px - w * py);
        if (s < 0.0) return len2(px, py) - r;
        if (px < w) return h - py;
        return len2(px - w, py - h);
    }
};

// sdTrapezoid — Shape.hpp:754-767; r1 = 1, r2 = 3, he = 2
template <>
struct ShapeFn<SH_TRAPEZOID> {
    static __device__ __forceinline__ double sdf(const ShapeParams &, double px, double py) {
        const double r1 = 1.0, r2 = 3.0, he = 2.0;
        const double k1x = r2, k1y = he;
        const double k2x = r2 - r1, k2y = 2.0 * he;
        px = fabs(px);
        double cax = smaxd(0.0, px - ((py < 0.0) ? r1 : r2));
        double cay = fabs(py) - he;
        double t = clipd(div_const((k1x - px) * k2x + (k1y - py) * k2y, KMISC[7], KMISC[8]), 0.0, 1.0);
        double cbx = px - k1x + k2x * t;
        double cby = py - k1y + k2y * t;
        double s = (cbx < 0.0 && cay < 0.0) ? -1.0 : 1.0;
        return s * sqrt(smind(cax * cax + cay * cay, cbx * cbx + cby * cby));
    }
};

// sdRhombus — Shape.hpp:809-826; b = (1, 4.5)
template <>
struct ShapeFn<SH_RHOMBUS> {
    static __device__ __forceinline__ double sdf(const ShapeParams &, double px, double py) {
        const double bx = 1.0, by = 4.5;
        px = fabs(px);
        py = fabs(py);
        double mx = bx - 2.0 * px, my = by - 2.0 * py;
        double h = clipd(div_const(mx * bx - my * by, KMISC[10], KMISC[11]), -1.0, 1.0);
        double hx = 0.5 * bx, hy = 0.5 * by;
        double dx = px - hx * (1.0 - h), dy = py - hy * (1.0 + h);
        double d = len2(dx, dy);
        double sign = signbit(px * by + py * bx - bx * by) ? -1.0 : 1.0;
        return d * sign;
    }
};

// sdHeart — Shape.hpp:939-952 (input / 4, output * 4); cst[0] = sqrt(2.0) / 4.0
template <>
struct ShapeFn<SH_HEART> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        px = px / 4.0;
        py = py / 4.0;
        px = fabs(px);
        if (py + px > 1.0) return 4 * (len2(px - 0.25, py - 0.75) - S.cst[0]);
        double ax = px - 0.0, ay = py - 1.0;
        double v1 = ax * ax + ay * ay;
        double t = smaxd(px + py, 0.0);
        double bx = px - 0.5 * t, by = py - 0.5 * t;
        double v2 = bx * bx + by * by;
        return 4 * (sqrt(smind(v1, v2)) * copysign(1.0, px - py));
    }
};

// sdRoundedX / bigX — Shape.hpp:988-994 (w = 3, r = 0.25) / 1024-1030 (w = 5, r = 0.25)
__device__ __forceinline__ double sd_roundedx_w(double px, double py, double w, double r) {
    double ax = fabs(px), ay = fabs(py);
    double m = (ax + ay > w) ? (w * 0.5) : (ax + ay) * 0.5;
    return len2(ax - m, ay - m) - r;
}
template <>
struct ShapeFn<SH_ROUNDEDX> {
    static __device__ __forceinline__ double sdf(const ShapeParams &, double px, double py) {
        return sd_roundedx_w(px, py, 3.0, 0.25);
    }
};
template <>
struct ShapeFn<SH_BIGX> {
    static __device__ __forceinline__ double sdf(const ShapeParams &, double px, double py) {
        return sd_roundedx_w(px, py, 5.0, 0.25);
    }
};

// sdRoundedCross — Shape.hpp:1062-1075; h = 1, input / 2, output * 2
template <>
struct ShapeFn<SH_ROUNDEDCROSS> {
    static __device__ __forceinline__ double sdf(const ShapeParams &, double px, double py) {
        const double h = 1.0;
        px = px / 2.0;
        py = py / 2.0;
        const double k = 0.5 * (h + 1.0 / h);
        double ax = fabs(px), ay = fabs(py);
        if (ax < 1.0 && ay < ax * (k - h) + h) return 2 * (k - len2(ax - 1.0, ay - k));
        double d1x = ax - 0.0, d1y = ay - h;
        double d2x = ax - 1.0, d2y = ay - 0.0;
        return 2 * sqrt(smind(d1x * d1x + d1y * d1y, d2x * d2x + d2y * d2y));
    }
};

// sdOrientedVesica — Shape.hpp:1115-1146; a = (2,4), b = (-2,-4), w = 0.8
// cst = (r, d, vx, vy) with r = 0.5*|b-a|, d = 0.5*(r*r - w*w)/w, v = (b-a)/r  (host-computed, same formulas)
template <>
struct ShapeFn<SH_VESICA> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        const double w = 0.8;
        const double r = S.cst[0], d = S.cst[1], vx = S.cst[2], vy = S.cst[3];
        const double cx = 0.5 * (-2.0 + 2.0), cy = 0.5 * (-4.0 + 4.0);
        px = px / 1.0;
        py = py / 1.0;
        double ux = px - cx, uy = py - cy;
        double qx = 0.5 * fabs(vy * ux + vx * uy);
        double qy = 0.5 * fabs(-vx * ux + vy * uy);
        double hx, hy, hz;
        if (r * qx < d * (qy - r)) {
            hx = 0.0; hy = r; hz = 0.0;
        } else {
            hx = -d; hy = 0.0; hz = d + w;
        }
        return 1.0 * (len2(qx - hx, qy - hy) - hz);
    }
};

// sdMoon — Shape.hpp:1202-1214; d = 0.8, ra = 3, rb = 2.4; cst = (a, b) host-computed
template <>
struct ShapeFn<SH_MOON> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double qx, double qy) {
        const double d = 0.8, ra = 3.0, rb = 2.4;
        const double a = S.cst[0], b = S.cst[1];
        qy = fabs(qy);
        bool cond = d * (qx * b - qy * a) > d * d * smaxd(b - qy, 0.0);
        double dist1 = len2(qx - a, qy - b);
        double dist2 = smaxd(len2(qx, qy) - ra, -len2(qx - d, qy - 0.0) + rb);
        return cond ? dist1 : dist2;
    }
};

// sdUnevenCapsule — Shape.hpp:531-543; r1 = 2, r2 = 1, h = 5; cst = (b, a) host-computed
template <>
struct ShapeFn<SH_UNEVENCAPSULE> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        const double r1 = 2.0, r2 = 1.0, h = 5.0;
        const double b = S.cst[0], a = S.cst[1];
        px = fabs(px);
        double k = px * (-b) + py * a;
        if (k < 0.0) return len2(px, py) - r1;
        if (k > a * h) return len2(px - 0.0, py - h) - r2;
        return px * a + py * b - r1;
    }
};

// Circle — Shape.hpp:476-480
template <>
struct ShapeFn<SH_CIRCLE> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double px, double py) {
        return len2(px, py) - S.radius;
    }
};

// Polygon fallback — Shape.hpp:1370-1400 (edge helpers), 1448-1476 (SDF), 1508-1534 (analytic gradient).
// Ignores trans/Rotate, like the reference (:1451).
struct PolyHit {
    double dis, cx, cy;
    int rs;
};
__device__ __forceinline__ PolyHit polygon_scan(const ShapeParams &S, double qx, double qy) {
    const double PI = 3.14159265358979323846;
    PolyHit H{1e9, 0.0, 0.0, 0};
#pragma unroll 1
    for (int i = 0; i < S.poly_n; ++i) {
        double sx = S.poly_sx[i], sy = S.poly_sy[i], ex = S.poly_ex[i], ey = S.poly_ey[i];
        double vx = ex - sx, vy = ey - sy;
        double wx = qx - sx, wy = qy - sy;
        double t = (wx * vx + wy * vy) / (vx * vx + vy * vy);
        if (t < 0.0) t = 0.0;
        else if (t > 1.0) t = 1.0;
        double cx = sx + t * vx, cy = sy + t * vy;
        double dis = len2(qx - cx, qy - cy);
        if (dis < H.dis) {
            H.dis = dis; H.cx = cx; H.cy = cy;
        }
        double ths = atan2_portable(sy - qy, sx - qx), the = atan2_portable(ey - qy, ex - qx);
        ths = (ths < 0.0) ? (ths + 2 * PI) : ths;
        the = (the < 0.0) ? (the + 2 * PI) : the;
        double d1 = fabs(ths - the);
        if (!(d1 < PI)) H.rs++;
    }
    return H;
}
template <>
struct ShapeFn<SH_POLYGON> {
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double qx, double qy) {
        PolyHit H = polygon_scan(S, qx, qy);
        return (H.rs % 2 == 0) ? H.dis : -H.dis;
    }
};

// Triangle-mesh functor — BasicShape::getonlySDF_igl (Shape.hpp:332-340): sdf = (1 - 2 w) * sqrt(d2).
//   w  = winding number: the reference asks igl::fast_winding_number (fast_winding_number.cpp:439-457), a float BVH whose
//        leaves evaluate the per-triangle solid angle exactly (UTsignedSolidAngleTri, FastWindingNumberForSoups.h:6071-6110)
//        and whose far clusters approximate the same sum; here: the sum itself, in double, same per-triangle formula.
//   d2 = squared distance to the closest triangle: igl::AABB::squared_distance (AABB.cpp:1130-1200) is a pruned minimum over
//        point_simplex_squared_distance (point_simplex_squared_distance.cpp:43-116, Ericson's closest point); here the plain
//        minimum over all faces — the same number.
// The query is (qx, qy, 0): the path zeroes the z of both the pose and the point (sw_manager.hpp:767,
// back_end_optimizer.hpp:791).  One pass over the faces does both sums; every lane of a warp reads the same face at the
// same time (uniform __ldg -> one L1 transaction per operand).
template <>
struct ShapeFn<SH_MESH> {
    static __device__ __forceinline__ double solid_angle(double ax, double ay, double az, double bx, double by, double bz,
                                                         double cx, double cy, double cz) {
        const double al = sqrt((ax * ax + ay * ay) + az * az);
        const double bl = sqrt((bx * bx + by * by) + bz * bz);
        const double cl = sqrt((cx * cx + cy * cy) + cz * cz);
        if (al == 0 || bl == 0 || cl == 0) return 0.0;
        const double ia = 1.0 / al, ib = 1.0 / bl, ic = 1.0 / cl;
        ax *= ia; ay *= ia; az *= ia;
        bx *= ib; by *= ib; bz *= ib;
        cx *= ic; cy *= ic; cz *= ic;
        const double ux = bx - ax, uy = by - ay, uz = bz - az;
        const double vx = cx - ax, vy = cy - ay, vz = cz - az;
        const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
        const double num = (ax * nx + ay * ny) + az * nz;
        if (num == 0) return 0.0;
        const double dab = (ax * bx + ay * by) + az * bz;
        const double dac = (ax * cx + ay * cy) + az * cz;
        const double dbc = (bx * cx + by * cy) + bz * cz;
        const double den = ((1.0 + dab) + dac) + dbc;
        return 2.0 * atan2_portable(num, den);
    }
    // squared distance from p to triangle (a, b, c); ap = p - a etc. are passed in (shared with the solid angle)
    static __device__ __forceinline__ double sqr_distance(double ax, double ay, double az, double bx, double by, double bz,
                                                          double cx, double cy, double cz, double px, double py, double pz) {
        const double abx = bx - ax, aby = by - ay, abz = bz - az;
        const double acx = cx - ax, acy = cy - ay, acz = cz - az;
        const double apx = px - ax, apy = py - ay, apz = pz - az;
        double qx = ax, qy = ay, qz = az;
        const double d1 = (abx * apx + aby * apy) + abz * apz;
        const double d2 = (acx * apx + acy * apy) + acz * apz;
        if (!(d1 <= 0.0 && d2 <= 0.0)) {
            const double bpx = px - bx, bpy = py - by, bpz = pz - bz;
            const double d3 = (abx * bpx + aby * bpy) + abz * bpz;
            const double d4 = (acx * bpx + acy * bpy) + acz * bpz;
            if (d3 >= 0.0 && d4 <= d3) {
                qx = bx; qy = by; qz = bz;
            } else {
                const double vc = d1 * d4 - d3 * d2;
                const bool a_ne_b = (ax != bx) || (ay != by) || (az != bz);
                if (a_ne_b && vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
                    const double v = d1 / (d1 - d3);
                    qx = ax + v * abx; qy = ay + v * aby; qz = az + v * abz;
                } else {
                    const double cpx = px - cx, cpy = py - cy, cpz = pz - cz;
                    const double d5 = (abx * cpx + aby * cpy) + abz * cpz;
                    const double d6 = (acx * cpx + acy * cpy) + acz * cpz;
                    if (d6 >= 0.0 && d5 <= d6) {
                        qx = cx; qy = cy; qz = cz;
                    } else {
                        const double vb = d5 * d2 - d1 * d6;
                        if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
                            const double w = d2 / (d2 - d6);
                            qx = ax + w * acx; qy = ay + w * acy; qz = az + w * acz;
                        } else {
                            const double va = d3 * d6 - d5 * d4;
                            if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
                                const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
                                qx = bx + w * (cx - bx); qy = by + w * (cy - by); qz = bz + w * (cz - bz);
                            } else {
                                const double denom = 1.0 / ((va + vb) + vc);
                                const double v = vb * denom, w = vc * denom;
                                qx = (ax + abx * v) + acx * w; qy = (ay + aby * v) + acy * w; qz = (az + abz * v) + acz * w;
                            }
                        }
                    }
                }
            }
        }
        const double ex = px - qx, ey = py - qy, ez = pz - qz;
        return (ex * ex + ey * ey) + ez * ez;
    }
    // One pass over the faces.  Record layout (kMeshStride doubles per face): a, b, c (9) and rmax = the largest distance
    // from vertex a to a point of the face, max(|ab|, |ac|), padded upwards on the host.  By the triangle inequality every
    // point of the face is at least |a - q| - rmax away from q, so when that bound (with a 1e-12 relative margin, four
    // orders above the rounding of the quantities involved) exceeds the best squared distance so far the face cannot
    // lower the minimum and its closest-point computation is skipped: a conservative prune, the minimum is unchanged
    // bit for bit (the oracle takes the plain minimum over all faces).  Skipped only when every lane of the warp agrees.
    static __device__ __forceinline__ double sdf(const ShapeParams &S, double qx, double qy) {
        const double PI = 3.14159265358979323846;
        const double qz = 0.0;
        double omega = 0.0, best = __longlong_as_double(0x7ff0000000000000LL);
        const double *t = S.mesh_tri;
#pragma unroll 1
        for (int f = 0; f < S.mesh_nf; ++f, t += kMeshStride) {
            const double ax = __ldg(t), ay = __ldg(t + 1), az = __ldg(t + 2);
            const double bx = __ldg(t + 3), by = __ldg(t + 4), bz = __ldg(t + 5);
            const double cx = __ldg(t + 6), cy = __ldg(t + 7), cz = __ldg(t + 8);
            const double rmax = __ldg(t + 9);
            const double rax = ax - qx, ray = ay - qy, raz = az - qz;
            omega += solid_angle(rax, ray, raz, bx - qx, by - qy, bz - qz, cx - qx, cy - qy, cz - qz);
            const double lb = sqrt((rax * rax + ray * ray) + raz * raz) - rmax;  // lower bound of the distance to the face
            const bool need = !(lb > 0.0 && lb * lb > best * (1.0 + 1e-12));
            if (__any_sync(__activemask(), need)) {
                const double d = sqr_distance(ax, ay, az, bx, by, bz, cx, cy, cz, qx, qy, qz);
                if (d < best) best = d;
            }
        }
        const double w = omega / (4.0 * PI);
        const double s = 1. - 2. * w;
        return s * sqrt(best);
    }
};

// ((pos_rel - trans) * Rotate).head(2): row-vector times matrix (Shape.hpp:281-294 and e.g. :586).
// With has_xform == 0 (trans = 0, Rotate = I) the product is the identity bit-for-bit and is skipped.
template <int SHAPE, bool XFORM>
__device__ __forceinline__ double shape_sdf(const ShapeParams &S, double rx, double ry) {
    if (SHAPE != SH_POLYGON && SHAPE != SH_MESH && XFORM) {
        double v0 = rx - S.trans[0], v1 = ry - S.trans[1];
        rx = v0 * S.rot[0] + v1 * S.rot[2];
        ry = v0 * S.rot[1] + v1 * S.rot[3];
    }
    return ShapeFn<SHAPE>::sdf(S, rx, ry);
}

// Circle overrides getonlyGrad1 (Shape.hpp:487-497): the normalised ((p - trans) * Rotate).head(2), no finite difference
template <bool XFORM>
__device__ __forceinline__ void circle_grad1(const ShapeParams &S, double rx, double ry, double &gx, double &gy) {
    if (XFORM) {
        double v0 = rx - S.trans[0], v1 = ry - S.trans[1];
        rx = v0 * S.rot[0] + v1 * S.rot[2];
        ry = v0 * S.rot[1] + v1 * S.rot[3];
    }
    double z = rx * rx + ry * ry;
    if (z > 0.0) {  // Eigen normalize(): divides by sqrt(squaredNorm)
        double n = sqrt(z);
        rx /= n; ry /= n;
    }
    gx = rx; gy = ry;
}

}  // namespace dev
}  // namespace svsdf
