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
// initShape (the yaw-indexed byte kernels) is PINNED to the reference's own BasicShape::initShape compiled from Shape.hpp
// (oracle/_ref/libref_path_*.so; tests/test_oracle_ref_pin.py::test_init_shape_kernels_match_reference_code, all 16 registry
// shapes, identical bytes).  kernelConv / checkKernelValue / the A* bookkeeping remain restatements (parity unpinned).  Pins:
// the two kernelConv variants of the reference must agree with each other, closed-form kernels of the Circle, and the
// byte-level layout shared with generateMapKernel2D (tests/test_oracle_frontend.py).
// The Polygon fallback is not covered: its rotated overload takes a Matrix2d and does not override the virtual the
// reference's initShape calls (Shape.hpp:1477 vs :267), i.e. the reference itself has no defined kernels for it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
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

// ---- the second half of the A* node test: the sub-swept-volume between father and child ----
// PCSmapManager::getPointsInAABB2D (map_manager/include/map_manager/PCSmap_manager.h:137-158) with projInMap (:126-133),
// GridMap3D::getGridIndex / getGridCubeCenter (map_manager/src/Gridmap3D.cpp:137-195): occupied cell centres in the box
struct MapGeom {
    double ox = 0, oy = 0, res = 1;  // boundary_xyzmin (x, y), grid_resolution; boundary_xyzmax = min + size * res
};
inline int grid_index_1d(double p, double lo, double res, int size) {
    int i = (int)std::floor((p - lo) / res);
    if (i < 0) i = 0;
    if (i >= size) i = size - 1;
    return i;
}
inline void points_in_aabb2d(const FrontMap &M, const MapGeom &G, double cx, double cy, double hbx, double hby, std::vector<double> &xy) {
    const double xmax = G.ox + M.X * G.res, ymax = G.oy + M.Y * G.res;
    double x1 = cx - hbx, y1 = cy - hby, x2 = cx + hbx, y2 = cy + hby;
    auto proj = [](double v, double lo, double hi) { if (v < lo) v = lo; if (v > hi) v = hi; return v; };
    x1 = proj(x1, G.ox, xmax); x2 = proj(x2, G.ox, xmax);
    y1 = proj(y1, G.oy, ymax); y2 = proj(y2, G.oy, ymax);
    const int i1 = grid_index_1d(x1, G.ox, G.res, M.X), i2 = grid_index_1d(x2, G.ox, G.res, M.X);
    const int j1 = grid_index_1d(y1, G.oy, G.res, M.Y), j2 = grid_index_1d(y2, G.oy, G.res, M.Y);
    for (int i = i1; i <= i2; i++)
        for (int j = j1; j <= j2; j++)
            if (M.occ[(size_t)i * M.Y + j]) {
                xy.push_back((i + 0.5) * G.res + G.ox);
                xy.push_back((j + 0.5) * G.res + G.oy);
            }
}

// SweptVolumeManager::checkSubSWCollision (sw_manager.hpp:1171-1210): every obstacle point must stay outside the shape
// while the pose moves linearly (yaw included) from father to child, sampled at kt = 0, 0.02, ... (accumulated) <= 1
inline bool check_sub_sw_collision(const Shape &S, const double father[3], const double child[3], const std::vector<double> &xy) {
    const double dt = 0.02;
    for (size_t i = 0; i + 1 < xy.size(); i += 2) {
        const double px = xy[i], py = xy[i + 1];
        double min_sdf = 1e9;
        for (double kt = 0.0; kt <= 1.0; kt += dt) {
            const double om = 1 - kt;
            const double lx = kt * child[0] + om * father[0];
            const double ly = kt * child[1] + om * father[1];
            const double yaw = kt * child[2] + om * father[2];
            double s, c;
            psc::sincos(yaw, s, c);
            const double d0 = px - lx, d1 = py - ly;
            const double rx = c * d0 + s * d1, ry = -s * d0 + c * d1;  // posEva2Rel (:521-526)
            const double f = shape_sdf(S, rx, ry, 0.0);
            if (f < min_sdf) min_sdf = f;
            if (min_sdf < 0) return false;
        }
    }
    return true;
}

// The neighbour loop of AstarPathSearcher::process (planner_algorithm/include/planner_algorithm/front_end_Astar.hpp:192-240)
// for one node (cell index, yaw): for each of the 9 cells (i, j) in -1..1:
//   cond = isIndexValid(vi) && !isIndexOccupiedFlate(vi, 0) && checkKernelValue(fy, cy, vi) && checkSubSWCollision(state1, state2, aabb)
// parts[n]: bit 0 valid and free, bit 1 kernel test, bit 2 sub-swept-volume test (each evaluated on its own, for tests);
// ok = all three; child_yaw = the yaw checkKernelValue chose (father's yaw when it failed).
inline void expand_node(const Shape &S, const ShapeKernels &K, const FrontMap &M, const MapGeom &G, int ix, int iy, double fy,
                        int kernel_size_world, uint8_t ok[9], double child_yaw[9], uint8_t parts[9]) {
    const double state1[3] = {(ix + 0.5) * G.res + G.ox, (iy + 0.5) * G.res + G.oy, fy};
    const double hb = (double)(kernel_size_world / 2 + 1);  // `kernel_size/2+1`, integer division (front_end_Astar.hpp:224)
    int n = 0;
    for (int i = -1; i < 2; i++)
        for (int j = -1; j < 2; j++, n++) {
            const int vx = ix + i, vy = iy + j;
            uint8_t p = 0;
            double cy = fy;
            const bool valid = M.valid(vx, vy);
            if (valid && !M.occ[(size_t)vx * M.Y + vy]) p |= 1;
            if (valid && check_kernel_value(K, M, fy, cy, vx, vy)) p |= 2;
            if (valid) {
                const double state2[3] = {(vx + 0.5) * G.res + G.ox, (vy + 0.5) * G.res + G.oy, cy};
                std::vector<double> pts;
                points_in_aabb2d(M, G, state2[0], state2[1], hb, hb, pts);
                if (check_sub_sw_collision(S, state1, state2, pts)) p |= 4;
            }
            parts[n] = p;
            ok[n] = (p == 7);
            child_yaw[n] = cy;
        }
}

// ---- AstarPathSearcher::AstarPathSearch / getPath (front_end_Astar.hpp:243-390), z = 0 layer ----
// Literal single-problem restatement: GridNode bookkeeping (id 0 unseen / 1 open / -1 closed, yaw fixed when a node is first
// seen), std::multimap open list (equal keys pop in insertion order), stale keys when an open node improves (the
// reference does not re-key it), closed nodes re-opened when improved, separate start-node object.
struct AstarResult {
    bool success = false;
    int expansions = 0;
    std::vector<double> path;  // x, y, yaw per node, start first
};
inline double astar_heuristic(int ax, int ay, int bx, int by) {  // getHeu (:165-183), dz = 0
    const double p = 1.0 / 1000;
    const int dx = std::abs(ax - bx), dy = std::abs(ay - by), dz = 0;
    const int dmin = std::min(dx, std::min(dy, dz));
    const int dmax = std::max(dx, std::max(dy, dz));
    const int dmid = dx + dy + dz - dmin - dmax;
    const double h = std::sqrt(3) * dmin + std::sqrt(2) * (dmid - dmin) + (dmax - dmid);
    return h * (1 + p);
}
inline AstarResult astar_search(const Shape &S, const ShapeKernels &K, const FrontMap &M, const MapGeom &G, const double start[2],
                                const double goal[2], int max_expansions = 1 << 30) {
    AstarResult R;
    const double xmax = G.ox + M.X * G.res, ymax = G.oy + M.Y * G.res;
    auto in_map = [&](const double p[2]) { return !(p[0] < G.ox || p[1] < G.oy || p[0] > xmax || p[1] > ymax); };
    if (!in_map(start) || !in_map(goal)) return R;
    const int sx = grid_index_1d(start[0], G.ox, G.res, M.X), sy = grid_index_1d(start[1], G.oy, G.res, M.Y);
    const int gx = grid_index_1d(goal[0], G.ox, G.res, M.X), gy = grid_index_1d(goal[1], G.oy, G.res, M.Y);
    struct Node { int id = 0; double g = 0, f = 0, yaw = 0; int father = -1; };  // father: node slot, -1 none
    const int NS = M.X * M.Y;                    // slot NS = the separate start-node object
    std::vector<Node> nodes((size_t)NS + 1);
    auto slot_xy = [&](int s, int &x, int &y) { if (s == NS) { x = sx; y = sy; } else { x = s / M.Y; y = s % M.Y; } };
    std::multimap<double, int> open;
    Node &st = nodes[NS];
    st.g = 0; st.f = astar_heuristic(sx, sy, gx, gy); st.id = 1; st.yaw = 0.0;
    open.insert({st.f, NS});
    nodes[(size_t)sx * M.Y + sy].id = 1; nodes[(size_t)sx * M.Y + sy].g = st.g; nodes[(size_t)sx * M.Y + sy].f = st.f;
    while (!open.empty()) {
        auto it = open.begin();
        const int cur = it->second;
        open.erase(it);
        nodes[cur].id = -1;
        int cx, cy_;
        slot_xy(cur, cx, cy_);
        if (cx == gx && cy_ == gy) {
            R.success = true;
            std::vector<int> chain;
            int p = cur;
            while (nodes[p].father != -1) { chain.push_back(p); p = nodes[p].father; }
            chain.push_back(p);
            for (auto q = chain.rbegin(); q != chain.rend(); ++q) {
                int x, y;
                slot_xy(*q, x, y);
                R.path.push_back((x + 0.5) * G.res + G.ox);
                R.path.push_back((y + 0.5) * G.res + G.oy);
                R.path.push_back(nodes[*q].yaw);
            }
            return R;
        }
        if (R.expansions >= max_expansions) return R;
        R.expansions++;
        uint8_t ok[9], parts[9];
        double cyaw[9];
        expand_node(S, K, M, G, cx, cy_, nodes[cur].yaw, K.kernel_size, ok, cyaw, parts);
        int n = 0;
        for (int i = -1; i < 2; i++)
            for (int j = -1; j < 2; j++, n++) {
                if (!ok[n]) continue;
                const int nx = cx + i, ny = cy_ + j, ns = nx * M.Y + ny;
                Node &nb = nodes[ns];
                if (nb.id == 0) nb.yaw = cyaw[n];  // AstarGetSucc :229-233
                const double ec = std::sqrt((double)(i * i + j * j));
                const double tg = ec + nodes[cur].g;
                if (nb.id == 0) {
                    nb.father = cur; nb.g = tg; nb.f = tg + astar_heuristic(nx, ny, gx, gy) + 0.0; nb.id = 1;
                    open.insert({nb.f, ns});
                } else if (nb.id == 1) {
                    if (tg < nb.g) { nb.father = cur; nb.g = tg; nb.f = tg + astar_heuristic(nx, ny, gx, gy) + 0.0; }  // key not updated
                } else {
                    if (tg < nb.g) {
                        nb.father = cur; nb.g = tg; nb.f = tg + astar_heuristic(nx, ny, gx, gy) + 0.0; nb.id = 1;
                        open.insert({nb.f, ns});
                    }
                }
            }
    }
    return R;
}

}  // namespace oracle
