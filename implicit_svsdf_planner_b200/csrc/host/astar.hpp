// host/astar.hpp — the reference's A* front end as a lock-step batch search over many start/goal problems.
//
// Restates AstarPathSearcher::AstarPathSearch / getPath (planner_algorithm/include/planner_algorithm/front_end_Astar.hpp:243-390,
// heuristic getHeu :165-183, z = 0 layer) for n independent problems on one map.  The search logic is the reference's, per
// problem: GridNode states (0 unseen / 1 open / -1 closed), a std::multimap open list (equal keys leave in insertion order),
// the yaw of a node fixed when it is first seen, open nodes improved in place WITHOUT re-keying, closed nodes re-opened when
// improved, a start-node object separate from the grid's node of the same cell.  What changes is the schedule: every
// iteration pops the best node of every unfinished problem and hands all of them to ONE call of `expand` (the neighbour
// loop of AstarGetSucc :192-240 — on the device: svsdf_front_expand), so the GPU sees thousands of nodes per launch
// instead of one.  `expand(m, node_ij, node_yaw, ok9, child_yaw9)` is a template parameter so that the host logic can be
// exercised without a GPU (tests/cpp/astar_host_main.cpp drives it with the CPU oracle's node test).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <unordered_map>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace svsdf {
namespace host {

struct AstarGrid {
    int X = 0, Y = 0;
    double ox = 0, oy = 0, res = 1;  // boundary_xyzmin (x, y), grid_resolution
};

inline int astar_grid_index(double p, double lo, double res, int size) {  // GridMap3D::getGridIndex, one axis (Gridmap3D.cpp:137-172)
    int i = (int)std::floor((p - lo) / res);
    if (i < 0) i = 0;
    if (i >= size) i = size - 1;
    return i;
}

inline double astar_heuristic(int ax, int ay, int bx, int by) {  // getHeu, dz = 0
    const double p = 1.0 / 1000;
    const int dx = std::abs(ax - bx), dy = std::abs(ay - by), dz = 0;
    const int dmin = std::min(dx, std::min(dy, dz));
    const int dmax = std::max(dx, std::max(dy, dz));
    const int dmid = dx + dy + dz - dmin - dmax;
    const double h = std::sqrt(3) * dmin + std::sqrt(2) * (dmid - dmin) + (dmax - dmid);
    return h * (1 + p);
}

struct AstarStats {
    int64_t expansions = 0;   // node expansions over all problems
    int64_t rounds = 0;       // expand() calls (lock-step iterations)
};

// paths_out: [n][max_path][3] (x, y, yaw; start first), len_out[n] = number of path nodes (0: no path, or longer than
// max_path), expansions_out[n] (may be null).  Returns 0, or the negative status an expand call returned.
template <class Expand>
int astar_batch(const AstarGrid &G, int n, const double *start_xy, const double *goal_xy, int max_path, double *paths_out, int32_t *len_out,
                int32_t *expansions_out, int64_t max_expansions_per_problem, Expand &&expand, AstarStats *stats = nullptr) {
    struct Node { int id = 0; double g = 0, f = 0, yaw = 0; int father = -1; };
    struct Problem {
        int sx = 0, sy = 0, gx = 0, gy = 0;
        bool done = false;
        int cur = -1;
        int64_t expansions = 0;
        std::unordered_map<int, Node> nodes;
        std::multimap<double, int> open;
    };
    const int NS = G.X * G.Y;  // slot of the separate start-node object
    const double xmax = G.ox + G.X * G.res, ymax = G.oy + G.Y * G.res;
    std::vector<Problem> P(n);
    for (int q = 0; q < n; ++q) {
        Problem &p = P[q];
        len_out[q] = 0;
        if (expansions_out) expansions_out[q] = 0;
        const double *s = start_xy + 2 * q, *g = goal_xy + 2 * q;
        auto in_map = [&](const double *v) { return !(v[0] < G.ox || v[1] < G.oy || v[0] > xmax || v[1] > ymax); };  // isInMap
        if (!in_map(s) || !in_map(g)) { p.done = true; continue; }
        p.sx = astar_grid_index(s[0], G.ox, G.res, G.X); p.sy = astar_grid_index(s[1], G.oy, G.res, G.Y);
        p.gx = astar_grid_index(g[0], G.ox, G.res, G.X); p.gy = astar_grid_index(g[1], G.oy, G.res, G.Y);
        Node &st = p.nodes[NS];
        st.g = 0; st.f = astar_heuristic(p.sx, p.sy, p.gx, p.gy); st.id = 1; st.yaw = 0.0;
        p.open.insert({st.f, NS});
        Node &cell = p.nodes[p.sx * G.Y + p.sy];
        cell.id = 1; cell.g = st.g; cell.f = st.f;
    }
    auto slot_xy = [&](const Problem &p, int s, int &x, int &y) { if (s == NS) { x = p.sx; y = p.sy; } else { x = s / G.Y; y = s % G.Y; } };
    std::vector<int> batch;
    std::vector<int32_t> ij;
    std::vector<double> yaw, cyaw;
    std::vector<unsigned char> ok;
    std::vector<unsigned char> wants(n, 0);   // problem q hands a node to this round's expand call
    std::vector<int32_t> cur_xy(2 * (size_t)n);
    std::vector<double> cur_yaw(n);
    // The searches are independent: the per-problem bookkeeping of a round (pop + goal test, then the neighbour updates)
    // runs on a few host threads when the batch is large; the order inside each problem is untouched.
#if defined(_OPENMP)
    const int host_threads = std::min(16, omp_get_max_threads());
#endif
    for (;;) {
#if defined(_OPENMP)
#pragma omp parallel for schedule(static) num_threads(host_threads) if (n >= 256)
#endif
        for (int q = 0; q < n; ++q) {
            Problem &p = P[q];
            wants[q] = 0;
            if (p.done) continue;
            if (p.open.empty()) { p.done = true; continue; }   // search failed
            auto it = p.open.begin();
            const int cur = it->second;
            p.open.erase(it);
            Node &c = p.nodes[cur];
            c.id = -1;
            int cx, cy;
            slot_xy(p, cur, cx, cy);
            if (cx == p.gx && cy == p.gy) {  // goal: getPath
                std::vector<int> chain;
                int s = cur;
                while (p.nodes[s].father != -1) { chain.push_back(s); s = p.nodes[s].father; }
                chain.push_back(s);
                if ((int)chain.size() <= max_path) {
                    double *out = paths_out + (size_t)q * max_path * 3;
                    int k = 0;
                    for (auto r = chain.rbegin(); r != chain.rend(); ++r, ++k) {
                        int x, y;
                        slot_xy(p, *r, x, y);
                        out[3 * k] = (x + 0.5) * G.res + G.ox;       // getGridCubeCenter
                        out[3 * k + 1] = (y + 0.5) * G.res + G.oy;
                        out[3 * k + 2] = p.nodes[*r].yaw;
                    }
                    len_out[q] = (int32_t)chain.size();
                }
                p.done = true;
                continue;
            }
            if (p.expansions >= max_expansions_per_problem) { p.done = true; continue; }
            p.expansions++;
            p.cur = cur;
            wants[q] = 1;
            cur_xy[2 * (size_t)q] = cx; cur_xy[2 * (size_t)q + 1] = cy;
            cur_yaw[q] = c.yaw;
        }
        batch.clear(); ij.clear(); yaw.clear();
        for (int q = 0; q < n; ++q)
            if (wants[q]) {
                batch.push_back(q);
                ij.push_back(cur_xy[2 * (size_t)q]); ij.push_back(cur_xy[2 * (size_t)q + 1]);
                yaw.push_back(cur_yaw[q]);
            }
        if (batch.empty()) break;
        const int m = (int)batch.size();
        ok.assign((size_t)9 * m, 0);
        cyaw.assign((size_t)9 * m, 0.0);
        const int rc = expand(m, ij.data(), yaw.data(), ok.data(), cyaw.data());
        if (rc != 0) return rc;
        if (stats) { stats->rounds++; stats->expansions += m; }
#if defined(_OPENMP)
#pragma omp parallel for schedule(static) num_threads(host_threads) if (m >= 256)
#endif
        for (int b = 0; b < m; ++b) {
            Problem &p = P[batch[b]];
            const int cur = p.cur;
            const int cx = ij[2 * b], cy = ij[2 * b + 1];
            const double gcur = p.nodes[cur].g;
            int k = 0;
            for (int i = -1; i < 2; i++)
                for (int j = -1; j < 2; j++, k++) {
                    if (!ok[(size_t)9 * b + k]) continue;
                    const int nx = cx + i, ny = cy + j, ns = nx * G.Y + ny;
                    Node &nb = p.nodes[ns];
                    if (nb.id == 0) nb.yaw = cyaw[(size_t)9 * b + k];  // AstarGetSucc :229-233
                    const double ec = std::sqrt((double)(i * i + j * j));
                    const double tg = ec + gcur;
                    if (nb.id == 0) {
                        nb.father = cur; nb.g = tg; nb.f = tg + astar_heuristic(nx, ny, p.gx, p.gy) + 0.0; nb.id = 1;
                        p.open.insert({nb.f, ns});
                    } else if (nb.id == 1) {
                        if (tg < nb.g) { nb.father = cur; nb.g = tg; nb.f = tg + astar_heuristic(nx, ny, p.gx, p.gy) + 0.0; }
                    } else if (tg < nb.g) {
                        nb.father = cur; nb.g = tg; nb.f = tg + astar_heuristic(nx, ny, p.gx, p.gy) + 0.0; nb.id = 1;
                        p.open.insert({nb.f, ns});
                    }
                }
        }
    }
    if (expansions_out)
        for (int q = 0; q < n; ++q) expansions_out[q] = (int32_t)P[q].expansions;
    return 0;
}

}  // namespace host
}  // namespace svsdf
