// oracle/shapes.hpp — TEST INFRASTRUCTURE ONLY (CPU restatement, never shipped, never on the product path).
//
// Plain-C++ (no Eigen) restatement of the reference's 2-D robot-shape SDF functors
//   /root/reference/src/utils/include/utils/Shape.hpp
// one function per shape class, same operation order as the reference source, double precision.
// PINNED to the reference's own classes: oracle/_ref/libref_path_*.so compiles the 16 registry classes + Circle + Polygon
// verbatim from Shape.hpp (oracle/ref_extract.py, oracle/ref_path_shim.cpp); tests/test_oracle_ref_pin.py requires
// getonlySDF / getonlyGrad1 of this file to equal them BIT FOR BIT (1e5 random points per shape live, 2 000 in the
// committed fixture tests/golden/ref_pin_shapes.npz, two body-frame pre-transforms).  The mesh functor's winding number is
// pinned against the reference's own compiled FWN code (oracle/_ref/libref_fwn.so, tests/golden/fwn_ref.npz).
// Further sanity pins in tests/: closed-form known answers, the Eikonal property, the reference's shapes/*.obj outlines.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "portable_sincos.hpp"
// The mesh functor's winding number is the reference's single-precision 4-way hierarchy with order-2 expansions.  That
// hierarchy (SAH build, 23 coefficient rows per node, traversal) is restated ONCE in this repository, in the product's host
// header below, and pinned there node by node / coefficient by coefficient / query by query against the reference's own
// code compiled into oracle/_ref/libref_fwn.so (tests/test_oracle_mesh.py) and against tests/golden/fwn_ref.npz; the oracle
// uses that host code (recursive traversal) as the checker of the device's explicit-stack traversal.
#include <memory>

#include "../implicit_svsdf_planner_b200/csrc/host/fwn_bvh.hpp"

namespace oracle {

enum ShapeId {
    SH_STAR = 0,
    SH_HORSESHOE,
    SH_PIE,
    SH_PIE2,
    SH_ARC,
    SH_TUNNEL,
    SH_CUTDISK,
    SH_TRAPEZOID,
    SH_RHOMBUS,
    SH_HEART,
    SH_ROUNDEDX,
    SH_BIGX,
    SH_ROUNDEDCROSS,
    SH_VESICA,
    SH_MOON,
    SH_UNEVENCAPSULE,
    SH_CIRCLE,
    SH_POLYGON,  // fallback for unknown names: sw_manager.hpp:363-372
    SH_MESH,     // triangle-mesh functor BasicShape::getonlySDF_igl (Shape.hpp:332-340), selected explicitly
    SH_COUNT
};

// Registry keys: sw_manager.hpp:187-235 (shapeConstructors). Unknown -> Polygon rect fallback.
inline int shape_id_from_name(const std::string &name) {
    static const char *names[] = {"star",        "sdHorseshoe", "sdPie",          "sdPie2",
                                  "sdArc",       "sdTunnel",    "sdCutDisk",      "sdTrapezoid",
                                  "sdRhombus",   "sdHeart",     "sdRoundedX",     "bigX",
                                  "sdRoundedCross", "sdOrientedVesica", "sdMoon", "sdUnevenCapsule"};
    for (int i = 0; i < 16; ++i)
        if (name == names[i]) return i;
    if (name == "Circle") return SH_CIRCLE;
    return SH_POLYGON;
}

struct Shape {
    int id = SH_STAR;
    // BasicShape ctor, Shape.hpp:281-294: trans=(p0,p1,0), Rotate=Rz(p2*PI/180)
    double trans[3] = {0, 0, 0};
    double Rot[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    // polygon fallback (edges start->end), Shape.hpp:1429-1446
    std::vector<double> poly_sx, poly_sy, poly_ex, poly_ey;
    double circle_radius = 1.0;
    // triangle soup of the mesh functor: 9 doubles per face (a, b, c), vertices already moved by the base-class
    // constructor's R v + trans (Shape.hpp:287-302)
    std::vector<double> mesh_tri;
    std::shared_ptr<svsdf::host::FwnBvh> fwn;  // igl::FastWindingNumberBVH of the transformed mesh (Shape.hpp:303-309)

    void set_poly_params(double p0, double p1, double p2_deg) {
        const double PI = 3.14159265358979323846;  // Shape.hpp:31
        trans[0] = p0;
        trans[1] = p1;
        trans[2] = 0.0;
        double yaw = (p2_deg * PI / 180.0);
        std::memset(Rot, 0, sizeof(Rot));
        Rot[0][0] = std::cos(yaw);
        Rot[0][1] = -std::sin(yaw);
        Rot[1][0] = std::sin(yaw);
        Rot[1][1] = std::cos(yaw);
        Rot[2][2] = 1;
    }
    void set_polygon(const double *xy, int n) {
        poly_sx.clear(); poly_sy.clear(); poly_ex.clear(); poly_ey.clear();
        for (int i = 0; i < n; ++i) {
            int j = (i + 1) % n;
            poly_sx.push_back(xy[2 * i]);
            poly_sy.push_back(xy[2 * i + 1]);
            poly_ex.push_back(xy[2 * j]);
            poly_ey.push_back(xy[2 * j + 1]);
        }
    }
    // BasicShape ctor, Shape.hpp:285-309: V <- hnormalized(homogeneous(V) * Trans^T), i.e. R v + trans per vertex;
    // V: nv x 3 row-major, F: nf x 3 (0-based)
    void set_mesh(const double *V, int nv, const int *F, int nf) {
        std::vector<double> Vt((size_t)nv * 3);
        for (int i = 0; i < nv; ++i) {
            const double *v = V + 3 * (size_t)i;
            for (int j = 0; j < 3; ++j) Vt[3 * (size_t)i + j] = ((v[0] * Rot[j][0] + v[1] * Rot[j][1]) + v[2] * Rot[j][2]) + trans[j];
        }
        mesh_tri.assign((size_t)nf * 9, 0.0);
        for (int f = 0; f < nf; ++f)
            for (int k = 0; k < 3; ++k)
                for (int j = 0; j < 3; ++j) mesh_tri[(size_t)f * 9 + 3 * k + j] = Vt[3 * (size_t)F[3 * f + k] + j];
        fwn = std::make_shared<svsdf::host::FwnBvh>();
        fwn->build(Vt.data(), nv, F, nf);  // igl::fast_winding_number(V.cast<float>(), F, 2, fwn_bvh), Shape.hpp:306-308
    }
    void set_default_rect() {  // sw_manager.hpp:365-369
        const double rect[8] = {6, -0.1, 6, 0.1, -6, 0.1, -6, -0.1};
        set_polygon(rect, 4);
    }
};

inline double clipd(double v, double lo, double hi) { return std::max(std::min(v, hi), lo); }

// ((pos_rel - trans) * Rotate).head(2): row-vector times matrix (Shape.hpp e.g. :586)
inline void pretransform(const Shape &S, double rx, double ry, double rz, double &px, double &py) {
    double v0 = rx - S.trans[0], v1 = ry - S.trans[1], v2 = rz - S.trans[2];
    px = v0 * S.Rot[0][0] + v1 * S.Rot[1][0] + v2 * S.Rot[2][0];
    py = v0 * S.Rot[0][1] + v1 * S.Rot[1][1] + v2 * S.Rot[2][1];
}

// ---- star: Shape.hpp:584-601 ----
inline double sd_star(double px, double py) {
    const double r = 2.8, rf = 0.6;
    const double k1x = 0.809016994375, k1y = -0.587785252292;
    const double k2x = -k1x, k2y = k1y;
    px = std::abs(px);
    double m = 2.0 * std::max(k1x * px + k1y * py, 0.0);
    px -= m * k1x;
    py -= m * k1y;
    m = 2.0 * std::max(k2x * px + k2y * py, 0.0);
    px -= m * k2x;
    py -= m * k2y;
    px = std::abs(px);
    py -= r;
    double bax = rf * (-k1y) - 0.0, bay = rf * k1x - 1.0;
    double h = clipd((px * bax + py * bay) / (bax * bax + bay * bay), 0.0, r);
    double dx = px - bax * h, dy = py - bay * h;
    return std::sqrt(dx * dx + dy * dy) * std::copysign(1.0, py * bax - px * bay);
}

// ---- sdHorseshoe: Shape.hpp:870-891; c=(cos 20.5, sin 20.5) radians (:855) ----
inline double sd_horseshoe(double px, double py) {
    const double r = 1.5;
    static const double cx = std::cos(20.5), cy = std::sin(20.5);
    const double wx = 1.55, wy = 0.20;
    px = std::abs(px);
    double l = std::sqrt(px * px + py * py);
    double qx = -cx * px + cy * py;
    double qy = cy * px + cx * py;
    double px0 = qx;
    if (px0 <= 0 && qy <= 0) qx = l * std::copysign(1.0, -cx);
    if (px0 <= 0) qy = l;
    qx = qx - wx;
    qy = std::abs(qy - r) - wy;
    double tx = std::max(qx, 0.0), ty = std::max(qy, 0.0);
    return std::sqrt(tx * tx + ty * ty) + std::min(0.0, std::max(qx, qy));
}

// ---- sdPie / sdPie2: Shape.hpp:1253-1260, 1294-1301; c=(cos 43, sin 43) / (cos 1, sin 1) ----
inline double sd_pie_c(double px, double py, double cx, double cy) {
    const double r = 3.0;
    px = std::abs(px);
    double l = std::sqrt(px * px + py * py) - r;
    double k = clipd(px * cx + py * cy, 0.0, r);
    double dx = px - cx * k, dy = py - cy * k;
    double m = std::sqrt(dx * dx + dy * dy);
    return std::max(l, m * std::copysign(1.0, cy * px - cx * py));
}
inline double sd_pie(double px, double py) {
    static const double cx = std::cos(43.0), cy = std::sin(43.0);
    return sd_pie_c(px, py, cx, cy);
}
inline double sd_pie2(double px, double py) {
    static const double cx = std::cos(1.0), cy = std::sin(1.0);
    return sd_pie_c(px, py, cx, cy);
}

// ---- sdArc: Shape.hpp:1334-1343; sc=(sin 20, cos 20), ra=2.3333, rb=0.5 ----
inline double sd_arc(double px, double py) {
    static const double scx = std::sin(20.0), scy = std::cos(20.0);
    const double ra = 2.3333, rb = 0.5;
    px = std::abs(px);
    bool cond = scy * px > scx * py;
    double ax = px - scx * ra, ay = py - scy * ra;
    double dist1 = std::sqrt(ax * ax + ay * ay);
    double dist2 = std::abs(std::sqrt(px * px + py * py) - ra);
    return (cond ? dist1 : dist2) - rb;
}

// ---- sdTunnel: Shape.hpp:642-658; wh=(2.5,1.5) ----
inline double sd_tunnel(double px, double py) {
    const double whx = 2.5, why = 1.5;
    px = std::abs(px);
    py = -py;
    double qx = px - whx, qy = py - why;
    double d1 = std::pow(std::max(qx, 0.0), 2) + qy * qy;
    qx = (py > 0.0) ? qx : std::sqrt(px * px + py * py) - whx;
    double d2 = std::pow(qx, 2) + std::pow(std::max(qy, 0.0), 2);
    double d = std::sqrt(std::min(d1, d2));
    return (std::max(qx, qy) < 0.0) ? -d : d;
}

// ---- sdCutDisk: Shape.hpp:698-711; r=5, h=2 ----
inline double sd_cutdisk(double px, double py) {
    const double r = 5.0, h = 2.0;
    const double w = std::sqrt(r * r - h * h);
    px = std::abs(px);
    double s = std::max((h - r) * px * px + w * w * (h + r - 2.0 * py), h * px - w * py);
    if (s < 0.0) return std::sqrt(px * px + py * py) - r;
    if (px < w) return h - py;
    double dx = px - w, dy = py - h;
    return std::sqrt(dx * dx + dy * dy);
}

// ---- sdTrapezoid: Shape.hpp:754-767; r1=1, r2=3, he=2 ----
inline double sd_trapezoid(double px, double py) {
    const double r1 = 1.0, r2 = 3.0, he = 2.0;
    const double k1x = r2, k1y = he;
    const double k2x = r2 - r1, k2y = 2.0 * he;
    px = std::abs(px);
    double cax = std::max(0.0, px - ((py < 0.0) ? r1 : r2));
    double cay = std::abs(py) - he;
    double t = clipd(((k1x - px) * k2x + (k1y - py) * k2y) / (k2x * k2x + k2y * k2y), 0.0, 1.0);
    double cbx = px - k1x + k2x * t;
    double cby = py - k1y + k2y * t;
    double s = (cbx < 0.0 && cay < 0.0) ? -1.0 : 1.0;
    return s * std::sqrt(std::min(cax * cax + cay * cay, cbx * cbx + cby * cby));
}

// ---- sdRhombus: Shape.hpp:809-826; b=(1,4.5) ----
inline double sd_rhombus(double px, double py) {
    const double bx = 1.0, by = 4.5;
    px = std::abs(px);
    py = std::abs(py);
    double mx = bx - 2.0 * px, my = by - 2.0 * py;
    double dotp = bx * bx + by * by;
    double h = clipd((mx * bx - my * by) / dotp, -1.0, 1.0);
    double hx = 0.5 * bx, hy = 0.5 * by;
    double dx = px - hx * (1.0 - h), dy = py - hy * (1.0 + h);
    double d = std::sqrt(dx * dx + dy * dy);
    double sign = std::signbit(px * by + py * bx - bx * by) ? -1.0 : 1.0;
    return d * sign;
}

// ---- sdHeart: Shape.hpp:939-952 (input /4, output *4) ----
inline double sd_heart(double px, double py) {
    px = px / 4.0;
    py = py / 4.0;
    px = std::abs(px);
    if (py + px > 1.0) {
        double ax = px - 0.25, ay = py - 0.75;
        return 4 * (std::sqrt(ax * ax + ay * ay) - std::sqrt(2.0) / 4.0);
    }
    double ax = px - 0.0, ay = py - 1.0;
    double v1 = ax * ax + ay * ay;
    double t = std::max(px + py, 0.0);
    double bx = px - 0.5 * t, by = py - 0.5 * t;
    double v2 = bx * bx + by * by;
    return 4 * (std::sqrt(std::min(v1, v2)) * std::copysign(1.0, px - py));
}

// ---- sdRoundedX / bigX: Shape.hpp:988-994, 1024-1030 ----
inline double sd_roundedx_w(double px, double py, double w, double r) {
    double ax = std::abs(px), ay = std::abs(py);
    double m = (ax + ay > w) ? (w * 0.5) : (ax + ay) * 0.5;
    double dx = ax - m, dy = ay - m;
    return std::sqrt(dx * dx + dy * dy) - r;
}
inline double sd_roundedx(double px, double py) { return sd_roundedx_w(px, py, 3.0, 0.25); }
inline double sd_bigx(double px, double py) { return sd_roundedx_w(px, py, 5.0, 0.25); }

// ---- sdRoundedCross: Shape.hpp:1062-1075; h=1, input /2, output *2 ----
inline double sd_roundedcross(double px, double py) {
    const double h = 1.0;
    px = px / 2.0;
    py = py / 2.0;
    double k = 0.5 * (h + 1.0 / h);
    double ax = std::abs(px), ay = std::abs(py);
    if (ax < 1.0 && ay < ax * (k - h) + h) {
        double dx = ax - 1.0, dy = ay - k;
        return 2 * (k - std::sqrt(dx * dx + dy * dy));
    }
    double d1x = ax - 0.0, d1y = ay - h;
    double d2x = ax - 1.0, d2y = ay - 0.0;
    return 2 * std::sqrt(std::min(d1x * d1x + d1y * d1y, d2x * d2x + d2y * d2y));
}

// ---- sdOrientedVesica: Shape.hpp:1115-1146; a=(2,4), b=(-2,-4), w=0.8 ----
inline double sd_vesica(double px, double py) {
    const double ax = 2, ay = 4, bx = -2, by = -4, w = 0.8;
    px = px / 1.0;
    py = py / 1.0;
    double bax = bx - ax, bay = by - ay;
    double r = 0.5 * std::sqrt(bax * bax + bay * bay);
    double d = 0.5 * (r * r - w * w) / w;
    double vx = bax / r, vy = bay / r;
    double cx = 0.5 * (bx + ax), cy = 0.5 * (by + ay);
    // rotation << v.y, v.x, -v.x, v.y  (row-major fill)
    double ux = px - cx, uy = py - cy;
    double qx = 0.5 * std::abs(vy * ux + vx * uy);
    double qy = 0.5 * std::abs(-vx * ux + vy * uy);
    // NOTE: reference takes cwiseAbs of (rotation*(p-c)) then multiplies by 0.5:
    //   q = 0.5 * (rotation * (p - c)).cwiseAbs()
    double hx, hy, hz;
    if (r * qx < d * (qy - r)) {
        hx = 0.0; hy = r; hz = 0.0;
    } else {
        hx = -d; hy = 0.0; hz = d + w;
    }
    double ex = qx - hx, ey = qy - hy;
    return 1.0 * (std::sqrt(ex * ex + ey * ey) - hz);
}

// ---- sdMoon: Shape.hpp:1202-1214; d=0.8, ra=3, rb=2.4 ----
inline double sd_moon(double qx, double qy) {
    const double d = 0.8, ra = 3.0, rb = 2.4;
    qy = std::abs(qy);
    double a = (ra * ra - rb * rb + d * d) / (2.0 * d);
    double b = std::sqrt(std::max(ra * ra - a * a, 0.0));
    bool cond = d * (qx * b - qy * a) > d * d * std::max(b - qy, 0.0);
    double e1x = qx - a, e1y = qy - b;
    double dist1 = std::sqrt(e1x * e1x + e1y * e1y);
    double e2x = qx - d, e2y = qy - 0.0;
    double dist2 = std::max(std::sqrt(qx * qx + qy * qy) - ra, -std::sqrt(e2x * e2x + e2y * e2y) + rb);
    return cond ? dist1 : dist2;
}

// ---- sdUnevenCapsule: Shape.hpp:531-543; r1=2, r2=1, h=5 ----
inline double sd_unevencapsule(double px, double py) {
    const double r1 = 2.0, r2 = 1.0, h = 5.0;
    px = std::abs(px);
    double b = (r1 - r2) / h;
    double a = std::sqrt(1.0 - b * b);
    double k = px * (-b) + py * a;
    if (k < 0.0) return std::sqrt(px * px + py * py) - r1;
    if (k > a * h) {
        double dx = px - 0.0, dy = py - h;
        return std::sqrt(dx * dx + dy * dy) - r2;
    }
    return px * a + py * b - r1;
}

// ---- Polygon fallback: Shape.hpp:1370-1400 (edge helpers), 1448-1476 (SDF) ----
struct PolyHit { double dis; double cx, cy; int rs; };
inline PolyHit polygon_scan(const Shape &S, double qx, double qy) {
    const double PI = 3.14159265358979323846;
    PolyHit H{1e9, 0, 0, 0};
    for (size_t i = 0; i < S.poly_sx.size(); ++i) {
        double sx = S.poly_sx[i], sy = S.poly_sy[i], ex = S.poly_ex[i], ey = S.poly_ey[i];
        // dis2Seg :1384-1399
        double vx = ex - sx, vy = ey - sy;
        double wx = qx - sx, wy = qy - sy;
        double t = (wx * vx + wy * vy) / (vx * vx + vy * vy);
        if (t < 0.0) t = 0.0;
        else if (t > 1.0) t = 1.0;
        double cx = sx + t * vx, cy = sy + t * vy;
        double ddx = qx - cx, ddy = qy - cy;
        double dis = std::sqrt(ddx * ddx + ddy * ddy);
        if (dis < H.dis) { H.dis = dis; H.cx = cx; H.cy = cy; }
        // isCrossRayOnXDir :1370-1383
        double s2x = sx - qx, s2y = sy - qy, e2x = ex - qx, e2y = ey - qy;
        double ths = psc::atan2(s2y, s2x), the = psc::atan2(e2y, e2x);
        ths = (ths < 0.0) ? (ths + 2 * PI) : ths;
        the = (the < 0.0) ? (the + 2 * PI) : the;
        double d1 = std::abs(ths - the);
        if (!(d1 < PI)) H.rs++;
    }
    return H;
}
inline double sd_polygon(const Shape &S, double qx, double qy) {
    PolyHit H = polygon_scan(S, qx, qy);
    return (H.rs % 2 == 0) ? H.dis : -H.dis;
}

// ---- Triangle-mesh functor: BasicShape::getonlySDF_igl, Shape.hpp:332-340 ----
//   sdf = (1 - 2 w) * sqrt(d2),  w = winding number of the mesh about the point, d2 = squared distance to the mesh.
// The reference gets w from igl::fast_winding_number (fast_winding_number.cpp:439-457 -> HDK UT_SolidAngle<float,float>,
// a float hierarchy whose leaves evaluate the per-triangle solid angle, UTsignedSolidAngleTri
// FastWindingNumberForSoups.h:6071-6110, and whose far clusters use an order-2 Taylor approximation of the same sum): a
// SINGLE-precision approximation, 2e-3 off the exact value on the reference's meshes, reproduced here through the shared host
// hierarchy (mesh_winding); the exact double sum is kept as mesh_winding_exact for the closed-form tests.  d2 comes from
// igl::AABB::squared_distance (AABB.cpp:1130-1200 -> point_simplex_squared_distance.cpp:43-116, Ericson's closest point
// on a triangle), restated as the plain minimum of the per-triangle squared distances (pruning does not change a minimum).
inline double tri_solid_angle(const double *t, double qx, double qy, double qz) {
    double ax = t[0] - qx, ay = t[1] - qy, az = t[2] - qz;
    double bx = t[3] - qx, by = t[4] - qy, bz = t[5] - qz;
    double cx = t[6] - qx, cy = t[7] - qy, cz = t[8] - qz;
    const double al = std::sqrt((ax * ax + ay * ay) + az * az);
    const double bl = std::sqrt((bx * bx + by * by) + bz * bz);
    const double cl = std::sqrt((cx * cx + cy * cy) + cz * cz);
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
    return 2.0 * psc::atan2(num, den);
}
inline double tri_sqr_distance(const double *t, double px, double py, double pz) {
    // ClosestBaryPtPointTriangle (Real-Time Collision Detection ch. 5; point_simplex_squared_distance.cpp:43-106)
    const double ax = t[0], ay = t[1], az = t[2], bx = t[3], by = t[4], bz = t[5], cx = t[6], cy = t[7], cz = t[8];
    const double abx = bx - ax, aby = by - ay, abz = bz - az;
    const double acx = cx - ax, acy = cy - ay, acz = cz - az;
    const double apx = px - ax, apy = py - ay, apz = pz - az;
    double qx, qy, qz;  // closest point
    const double d1 = (abx * apx + aby * apy) + abz * apz;
    const double d2 = (acx * apx + acy * apy) + acz * apz;
    bool done = false;
    if (d1 <= 0.0 && d2 <= 0.0) { qx = ax; qy = ay; qz = az; done = true; }
    double d3 = 0, d4 = 0, d5 = 0, d6 = 0, vc = 0, vb = 0;
    if (!done) {
        const double bpx = px - bx, bpy = py - by, bpz = pz - bz;
        d3 = (abx * bpx + aby * bpy) + abz * bpz;
        d4 = (acx * bpx + acy * bpy) + acz * bpz;
        if (d3 >= 0.0 && d4 <= d3) { qx = bx; qy = by; qz = bz; done = true; }
    }
    if (!done) {
        vc = d1 * d4 - d3 * d2;
        const bool a_ne_b = (ax != bx) || (ay != by) || (az != bz);
        if (a_ne_b && vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
            const double v = d1 / (d1 - d3);
            qx = ax + v * abx; qy = ay + v * aby; qz = az + v * abz; done = true;
        }
    }
    if (!done) {
        const double cpx = px - cx, cpy = py - cy, cpz = pz - cz;
        d5 = (abx * cpx + aby * cpy) + abz * cpz;
        d6 = (acx * cpx + acy * cpy) + acz * cpz;
        if (d6 >= 0.0 && d5 <= d6) { qx = cx; qy = cy; qz = cz; done = true; }
    }
    if (!done) {
        vb = d5 * d2 - d1 * d6;
        if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
            const double w = d2 / (d2 - d6);
            qx = ax + w * acx; qy = ay + w * acy; qz = az + w * acz; done = true;
        }
    }
    if (!done) {
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
    const double ex = px - qx, ey = py - qy, ez = pz - qz;
    return (ex * ex + ey * ey) + ez * ez;
}
// what the reference computes: the float hierarchy, accuracy_scale 2 (Shape.hpp:337)
inline double mesh_winding(const Shape &S, double qx, double qy, double qz) {
    const double p[3] = {qx, qy, qz};
    return S.fwn->winding_number(p, 2.0f);
}
// the quantity that hierarchy approximates: the exact double-precision sum over all faces (kept for the known-answer tests)
inline double mesh_winding_exact(const Shape &S, double qx, double qy, double qz) {
    const double PI = 3.14159265358979323846;
    double omega = 0.0;
    const size_t nf = S.mesh_tri.size() / 9;
    for (size_t f = 0; f < nf; ++f) omega += tri_solid_angle(&S.mesh_tri[9 * f], qx, qy, qz);
    return omega / (4.0 * PI);  // fast_winding_number.cpp:453
}
inline double mesh_sqr_distance(const Shape &S, double qx, double qy, double qz) {
    double best = std::numeric_limits<double>::infinity();
    const size_t nf = S.mesh_tri.size() / 9;
    for (size_t f = 0; f < nf; ++f) {
        const double d = tri_sqr_distance(&S.mesh_tri[9 * f], qx, qy, qz);
        if (d < best) best = d;
    }
    return best;
}
inline double sd_mesh(const Shape &S, double qx, double qy, double qz) {
    const double w = mesh_winding(S, qx, qy, qz);
    const double s = 1. - 2. * w;
    return s * std::sqrt(mesh_sqr_distance(S, qx, qy, qz));
}

inline double sd_mesh_exact(const Shape &S, double qx, double qy, double qz) {
    const double w = mesh_winding_exact(S, qx, qy, qz);
    const double s = 1. - 2. * w;
    return s * std::sqrt(mesh_sqr_distance(S, qx, qy, qz));
}

// BasicShape::getonlySDF(pos_rel) dispatch (virtual call in the reference, Shape.hpp:266)
inline double shape_sdf(const Shape &S, double rx, double ry, double rz) {
    if (S.id == SH_POLYGON) return sd_polygon(S, rx, ry);  // ignores trans/Rotate (:1451)
    if (S.id == SH_MESH) return sd_mesh(S, rx, ry, rz);    // the vertices carry the transform (:296-302)
    double px, py;
    pretransform(S, rx, ry, rz, px, py);
    switch (S.id) {
        case SH_STAR: return sd_star(px, py);
        case SH_HORSESHOE: return sd_horseshoe(px, py);
        case SH_PIE: return sd_pie(px, py);
        case SH_PIE2: return sd_pie2(px, py);
        case SH_ARC: return sd_arc(px, py);
        case SH_TUNNEL: return sd_tunnel(px, py);
        case SH_CUTDISK: return sd_cutdisk(px, py);
        case SH_TRAPEZOID: return sd_trapezoid(px, py);
        case SH_RHOMBUS: return sd_rhombus(px, py);
        case SH_HEART: return sd_heart(px, py);
        case SH_ROUNDEDX: return sd_roundedx(px, py);
        case SH_BIGX: return sd_bigx(px, py);
        case SH_ROUNDEDCROSS: return sd_roundedcross(px, py);
        case SH_VESICA: return sd_vesica(px, py);
        case SH_MOON: return sd_moon(px, py);
        case SH_UNEVENCAPSULE: return sd_unevencapsule(px, py);
        case SH_CIRCLE: return std::sqrt(px * px + py * py) - S.circle_radius;  // :476-480
        default: return 1e9;
    }
}

// getonlyGrad1: DEFINE_USEFUL_FUNCTION macro, Shape.hpp:35-53 (central FD, dx = 1e-6, in the body
// frame, BEFORE the shape pre-transform); Polygon overrides it analytically (:1508-1534).
inline void shape_grad1(const Shape &S, double rx, double ry, double rz, double g[3]) {
    if (S.id == SH_POLYGON) {
        PolyHit H = polygon_scan(S, rx, ry);
        double vx = rx - H.cx, vy = ry - H.cy, vz = rz - rz;
        double z = vx * vx + vy * vy + vz * vz;
        if (z > 0.0) {  // Eigen normalized()
            double n = std::sqrt(z);
            vx /= n; vy /= n; vz /= n;
        }
        if (H.rs % 2 != 0) { vx = -vx; vy = -vy; vz = -vz; }
        g[0] = vx; g[1] = vy; g[2] = vz;
        return;
    }
    if (S.id == SH_CIRCLE) {  // Circle overrides getonlyGrad1 (Shape.hpp:487-497): normalised ((p - trans) * Rotate).head(2)
        double px, py;
        pretransform(S, rx, ry, rz, px, py);
        double z = px * px + py * py;
        if (z > 0.0) {  // Eigen normalize()
            double n = std::sqrt(z);
            px /= n; py /= n;
        }
        g[0] = px; g[1] = py; g[2] = 0.0;
        return;
    }
    double dx = 0.000001;
    double t0 = rx, t1 = ry;
    t0 -= dx;
    double sdfold = shape_sdf(S, t0, t1, rz);
    t0 += 2 * dx;
    double gradx = shape_sdf(S, t0, t1, rz) - sdfold;
    t0 = rx;
    t1 -= dx;
    sdfold = shape_sdf(S, t0, t1, rz);
    t1 += 2 * dx;
    double grady = shape_sdf(S, t0, t1, rz) - sdfold;
    g[0] = gradx / (2 * dx);
    g[1] = grady / (2 * dx);
    g[2] = 0 / (2 * dx);
}

}  // namespace oracle
