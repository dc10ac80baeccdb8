// CPU check of the product's lock-step batch A* (implicit_svsdf_planner_b200/csrc/host/astar.hpp): the node test is supplied by the
// CPU oracle here (tests may use the oracle; the product passes svsdf_front_expand), and every problem's path and expansion
// count must equal the oracle's literal single-problem restatement of AstarPathSearch.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../implicit_svsdf_planner_b200/csrc/host/astar.hpp"
#include "../../implicit_svsdf_planner_b200/csrc/host/astar_flat.hpp"
#include "../../oracle/frontend_oracle.hpp"

int main(int argc, char **argv) {
    using namespace oracle;
    const int X = 48, Y = 41, ks = 17, K = 18, n = argc > 1 ? std::atoi(argv[1]) : 24;
    const double res = 1.0, ox = -7.25, oy = 3.5;
    std::mt19937 rng(12345);
    std::vector<uint8_t> occ((size_t)X * Y, 0);
    for (auto &c : occ) c = (rng() % 1000) < 6;
    Shape S;
    S.id = shape_id_from_name("star");
    ShapeKernels SK = init_shape_kernels(S, ks, K, res, 0.0);
    FrontMap M;
    M.build(occ.data(), X, Y, ks);
    MapGeom G;
    G.ox = ox; G.oy = oy; G.res = res;
    std::vector<double> st(2 * n), go(2 * n);
    std::uniform_real_distribution<double> ux(ox - 1.0, ox + X * res + 1.0), uy(oy - 1.0, oy + Y * res + 1.0);  // some outside the map
    for (int q = 0; q < n; ++q) { st[2 * q] = ux(rng); st[2 * q + 1] = uy(rng); go[2 * q] = ux(rng); go[2 * q + 1] = uy(rng); }
    st[0] = ox + 20.5; st[1] = oy + 20.5; go[0] = ox + 20.75; go[1] = oy + 20.25;  // start and goal in the same cell: a one-node path
    const int max_path = 256;
    std::vector<double> paths((size_t)n * max_path * 3, 0.0);
    std::vector<int32_t> len(n), ex(n);
    svsdf::host::AstarGrid AG;
    AG.X = X; AG.Y = Y; AG.ox = ox; AG.oy = oy; AG.res = res;
    svsdf::host::AstarStats stats;
    auto expand = [&](int m, const int32_t *ij, const double *yaw, unsigned char *ok, double *cyaw) {
        for (int b = 0; b < m; ++b) {
            uint8_t parts[9];
            expand_node(S, SK, M, G, ij[2 * b], ij[2 * b + 1], yaw[b], ks, ok + 9 * b, cyaw + 9 * b, parts);
        }
        return 0;
    };
    int rc = svsdf::host::astar_batch(AG, n, st.data(), go.data(), max_path, paths.data(), len.data(), ex.data(), 1 << 30, expand, &stats);
    if (rc != 0) { std::printf("FAIL rc %d\n", rc); return 1; }
    int found = 0, bad = 0;
    for (int q = 0; q < n; ++q) {
        AstarResult R = astar_search(S, SK, M, G, &st[2 * q], &go[2 * q]);
        const int want_len = R.success ? (int)(R.path.size() / 3) : 0;
        if (want_len != len[q] || R.expansions != ex[q]) { bad++; std::printf("problem %d: len %d vs %d, expansions %d vs %d\n", q, len[q], want_len, ex[q], R.expansions); continue; }
        if (want_len && std::memcmp(R.path.data(), &paths[(size_t)q * max_path * 3], R.path.size() * sizeof(double)) != 0) { bad++; std::printf("problem %d: path differs\n", q); }
        found += want_len > 0;
    }
    if (len[0] != 1) { bad++; std::printf("same-cell problem: len %d\n", len[0]); }
    // the container-free bookkeeping (host/astar_flat.hpp: flat arrays + a binary heap ordered by (f, insertion number); the code a
    // device-side search would run): one search at a time, same node test, must give the same paths and expansion counts
    {
        const int NS = X * Y;
        std::vector<int8_t> id(NS + 1);
        std::vector<double> g(NS + 1), f(NS + 1), yw(NS + 1);
        std::vector<int32_t> father(NS + 1);
        std::vector<svsdf::host::FlatHeapEntry> heap(4 * (size_t)(NS + 1));
        const double xmax = ox + X * res, ymax = oy + Y * res;
        int flat_bad = 0;
        size_t max_heap = 0;
        for (int q = 0; q < n; ++q) {
            const double *s = &st[2 * q], *gq = &go[2 * q];
            int flen = 0, fex = 0;
            std::vector<double> fpath((size_t)max_path * 3, 0.0);
            auto in_map = [&](const double *v) { return !(v[0] < ox || v[1] < oy || v[0] > xmax || v[1] > ymax); };
            if (in_map(s) && in_map(gq)) {
                std::fill(id.begin(), id.end(), 0);
                std::fill(father.begin(), father.end(), -1);
                svsdf::host::FlatSearch F;
                F.X = X; F.Y = Y; F.NS = NS; F.id = id.data(); F.g = g.data(); F.f = f.data(); F.yaw = yw.data(); F.father = father.data();
                F.heap = heap.data(); F.heap_cap = (int)heap.size();
                F.begin(svsdf::host::astar_grid_index(s[0], ox, res, X), svsdf::host::astar_grid_index(s[1], oy, res, Y),
                        svsdf::host::astar_grid_index(gq[0], ox, res, X), svsdf::host::astar_grid_index(gq[1], oy, res, Y));
                int cx, cy;
                double cyaw;
                while (F.pop_next(1LL << 30, cx, cy, cyaw)) {
                    unsigned char ok9[9];
                    double yaw9[9];
                    uint8_t parts[9];
                    expand_node(S, SK, M, G, cx, cy, cyaw, ks, ok9, yaw9, parts);
                    F.apply(cx, cy, ok9, yaw9);
                    max_heap = std::max(max_heap, (size_t)F.heap_n);
                }
                if (F.status == 1) {
                    flen = F.path(ox, oy, res, max_path, fpath.data());
                    if (flen > max_path) flen = 0;
                }
                if (F.status < 0) { flat_bad++; std::printf("flat problem %d: heap overflow\n", q); }
                fex = (int)F.expansions;
            }
            if (flen != len[q] || fex != ex[q]) { flat_bad++; std::printf("flat problem %d: len %d vs %d, expansions %d vs %d\n", q, flen, len[q], fex, ex[q]); continue; }
            if (flen && std::memcmp(fpath.data(), &paths[(size_t)q * max_path * 3], (size_t)flen * 3 * sizeof(double)) != 0) { flat_bad++; std::printf("flat problem %d: path differs\n", q); }
        }
        std::printf("flat bookkeeping: %s (largest open list %zu entries of capacity %zu)\n", flat_bad ? "FAIL" : "identical", max_heap, heap.size());
        bad += flat_bad;
    }
    std::printf("%s problems %d found %d rounds %lld expansions %lld\n", bad ? "FAIL" : "OK", n, found, (long long)stats.rounds, (long long)stats.expansions);
    return bad ? 1 : 0;
}
