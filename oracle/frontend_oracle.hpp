// oracle/frontend_oracle.hpp — TEST INFRASTRUCTURE ONLY (CPU restatement, never shipped, never on the product path).
//
// Restates the collision-kernel machinery of the reference's A* front end, paths relative to
// /root/reference/src:
//   BasicShape::initShape                 utils/include/utils/Shape.hpp:386-430   yaw-indexed occupancy kernels of the shape
//   shapeKernel / byteShapeKernel         Shape.hpp:100-218                       bool kernel, MSB-first byte kernel
//   getonlySDF(pos_rel, R_obj)            Shape.hpp:481-485, 545-556, ... (every analytic class): ((p - trans) Rotate R_obj).head(2)
//   SweptVolumeManager::kernelConv        swept_volume/include/swept_volume/sw_manager.hpp:1033-1096 (bool and byte variants)
//   visit_kernels_by_distance             sw_manager.hpp:1099-1156
//   checkKernelValue                      sw_manager.hpp:1158-1169   (`#define pi 3.1415926536`, :20)
// PARITY UNPINNED (no reference tests / golden vectors for this path; the reference cannot be compiled here).  Pins:
// the two kernelConv variants of the reference must agree with each other, closed-form kernels of the Circle, and the
// byte-level layout shared with generateMapKernel2D (tests/test_oracle_frontend.py).
// The Polygon fallback is not covered: its rotated overload takes a Matrix2d and does not override the virtual the
// reference's initShape calls (Shape.hpp:1477 vs :267), i.e. the reference itself has no defined kernels for it.
#pragma once
#include <cstdint>
#include <queue>
#include <vector>

#include "shapes.hpp"

namespace oracle {

static const uint8_t kOrMask[8] = {0x80, 0x40, 0x20, 0x10, 0x08, 0x04, 0x02, 0x01};  // PCSmap_manager.h:32 / Shape.hpp or_mask

// shape functor dispatch after the rotated pre-transform (px, py already in the functor's frame)
inline double shape_value_2d(const Shape &S, double px, double py) {
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
        case SH_CIRCLE: return std::sqrt(px * px + py * py) - S.circle_radius;
        default: return 1e9;
    }
}

// getonlySDF(pos_rel, R_obj) with R_obj = AngleAxisd(yaw, UnitZ): ((pos_rel - trans) * Rotate * R_obj).head(2).
// Row vector times matrix, left to right; the z row/column of both matrices is (0, 0, 1) and pos z = 0, so only the
// 2 x 2 blocks contribute (the omitted terms are exact zeros).
inline double shape_sdf_rotated(const Shape &S, double x, double y, double c, double s) {
    const double v0 = x - S.trans[0], v1 = y - S.trans[1];
    const double w0 = v0 * S.Rot[0][0] + v1 * S.Rot[1][0];
    const double w1 = v0 * S.Rot[0][1] + v1 * S.Rot[1][1];
    // R_obj = [[c, -s], [s, c]]
    const double u0 = w0 * c + w1 * s;
    const double u1 = w0 * (-s) + w1 * c;
    return shape_value_2d(S, u0, u1);
}

struct ShapeKernels {
    int kernel_size = 0, kernel_count = 0;
    std::vector<double> yaw;                 // [kernel_count]
    std::vector<uint8_t> cells;              // bool kernels  [kernel_count][ks][ks]
    std::vector<uint8_t> bytes;              // byte kernels  [kernel_count][ks][(ks + 7) / 8], MSB first
    int bytes_per_row() const { return (kernel_size + 7) / 8; }
};

// initShape (Shape.hpp:386-430) + generateByteKernel (:194-216)
inline ShapeKernels init_shape_kernels(const Shape &S, int kernel_size, int kernel_count, double kernelresu, double front_end_safeh) {
    const double PI = 3.14159265358979323846;  // Shape.hpp:31
    ShapeKernels K;
    K.kernel_size = kernel_size;
    K.kernel_count = kernel_count;
    const int size_side = (int)(0.5 * (kernel_size - 1));
    const double safemargin = std::max(front_end_safeh, kernelresu / 2);
    const double yaw_res = 2 * PI / kernel_count;
    K.yaw.assign(kernel_count, 0.0);
    K.cells.assign((size_t)kernel_count * kernel_size * kernel_size, 0);
    const int bpr = K.bytes_per_row();
    K.bytes.assign((size_t)kernel_count * kernel_size * bpr, 0);
    int ind = 0;
    // the reference's loop is `for (yaw = -PI; yaw < PI; yaw += yaw_res, ind++)`; it can run one time too many when the
    // accumulated sum stays below PI (writing past its arrays) — the restatement stops at kernel_count
    for (double yaw = -PI; yaw < PI && ind < kernel_count; yaw += yaw_res, ind++) {
        K.yaw[ind] = yaw;
        double s, c;
        psc::sincos(yaw, s, c);
        for (int a = 0; a < kernel_size; a++)
            for (int b = 0; b < kernel_size; b++) {
                const double x = kernelresu * a - size_side * kernelresu;
                const double y = kernelresu * b - size_side * kernelresu;
                const double sdf = shape_sdf_rotated(S, x, y, c, s);
                if (sdf <= safemargin) {
                    K.cells[((size_t)ind * kernel_size + a) * kernel_size + b] = 1;
                    K.bytes[((size_t)ind * kernel_size + a) * bpr + b / 8] |= kOrMask[b % 8];
                }
            }
    }
    return K;
}

// The occupancy map as the reference's front end sees it: plain grid (for the bool variant) and the inflated,
// byte-packed map kernel of generateMapKernel2D (PCSmap_manager.h:81-108).
struct FrontMap {
    int X = 0, Y = 0, h = 0, row_bytes = 0;
    std::vector<uint8_t> occ;     // [X][Y] 0/1
    std::vector<uint8_t> kernel;  // [(X + 2h)][row_bytes]
    void build(const uint8_t *occ_xy, int X_, int Y_, int kernel_size) {
        X = X_; Y = Y_; h = (kernel_size - 1) / 2;
        row_bytes = (Y + 2 * h + 7) / 8;
        occ.assign(occ_xy, occ_xy + (size_t)X * Y);
        kernel.assign((size_t)(X + 2 * h) * row_bytes, 0);
        for (int i = 0; i < X; ++i)
            for (int j = 0; j < Y; ++j)
                if (occ[(size_t)i * Y + j]) kernel[(size_t)(i + h) * row_bytes + (j + h) / 8] |= kOrMask[(j + h) % 8];
    }
    bool valid(int i, int j) const { return i >= 0 && i < X && j >= 0 && j < Y; }
};

// kernelConv<false> (sw_manager.hpp:1043-1066): true = no occupied cell under the shape kernel
inline bool kernel_conv_bool(const ShapeKernels &K, const FrontMap &M, int kernel_i, int ind_x, int ind_y) {
    const int ks = K.kernel_size, side = (ks - 1) / 2;
    for (int off_x = -side; off_x <= side; off_x++)
        for (int off_y = -side; off_y <= side; off_y++) {
            if (!M.valid(off_x + ind_x, off_y + ind_y)) continue;
            const int a = off_x + side, b = off_y + side;
            if (!K.cells[((size_t)kernel_i * ks + a) * ks + b]) continue;
            if (!M.occ[(size_t)(off_x + ind_x) * M.Y + (off_y + ind_y)]) continue;
            return false;
        }
    return true;
}

// kernelConv<true> (sw_manager.hpp:1068-1095): byte-AND of the byte kernel with the inflated map kernel window.
// Reads past the end of a row only ever feed bits the byte kernel masks out; they are guarded here instead of left to
// chance.
inline bool kernel_conv_byte(const ShapeKernels &K, const FrontMap &M, int kernel_i, int ind_x, int ind_y) {
    const int ks = K.kernel_size, bpr = K.bytes_per_row();
    const size_t total = M.kernel.size();
    for (int i = 0; i < ks; i++) {
        const size_t start = (size_t)(ind_x + i) * M.row_bytes + (ind_y / 8);
        const int off = ind_y % 8;
        for (int j = 0; j < bpr; j++) {
            const uint8_t m0 = (start + j < total) ? M.kernel[start + j] : 0;
            const uint8_t m1 = (start + j + 1 < total) ? M.kernel[start + j + 1] : 0;
            const uint8_t block = (uint8_t)((m0 << off) | (m1 >> (8 - off)));
            if (K.bytes[((size_t)kernel_i * ks + i) * bpr + j] & block) return false;
        }
    }
    return true;
}

// visit_kernels_by_distance (sw_manager.hpp:1099-1156): breadth-first over the yaw ring from start_x, at most
// maxdeepth + 1 kernels are tried
inline bool visit_kernels_by_distance(const ShapeKernels &K, const FrontMap &M, int &returni, int start_x, int ind_x, int ind_y,
                                      int maxdeepth = 10) {
    const int count = K.kernel_count;
    std::vector<bool> visited(count, false);
    std::queue<int> q;
    q.push(start_x);
    visited[start_x] = true;
    int deep = 0;
    while (!q.empty()) {
        deep++;
        const int x = q.front();
        q.pop();
        if (kernel_conv_byte(K, M, x, ind_x, ind_y)) {
            returni = x;
            return true;
        }
        for (int dir : {-1, 1}) {
            int nx = x + dir;
            if (nx < 0) nx = count - 1;
            if (nx >= count) nx = 0;
            if (visited[nx]) continue;
            visited[nx] = true;
            q.push(nx);
        }
        if (deep > maxdeepth) return false;
    }
    return false;
}

// checkKernelValue (sw_manager.hpp:1158-1169).  father_i == kernel_count (father_yaw at +pi) indexes past the
// reference's arrays; it is clamped to the last kernel here.
inline bool check_kernel_value(const ShapeKernels &K, const FrontMap &M, double father_yaw, double &child_yaw, int ind_x, int ind_y) {
    const double pi = 3.1415926536;  // sw_manager.hpp:20
    int father_i = int(K.kernel_count * ((father_yaw + pi) / (2 * pi)));
    if (father_i < 0) father_i = 0;
    if (father_i >= K.kernel_count) father_i = K.kernel_count - 1;
    int ret_i = father_i;
    if (visit_kernels_by_distance(K, M, ret_i, father_i, ind_x, ind_y)) {
        child_yaw = 2 * pi * (ret_i) / K.kernel_count - pi;
        return true;
    }
    return false;
}

}  // namespace oracle
