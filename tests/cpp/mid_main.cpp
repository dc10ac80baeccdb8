// Drives the mid end through the C++ mirror (include/svsdf.hpp: svsdf::OriTraj) the way plan_manager.cpp:176-192 drives the original:
// reads a problem (N, initS, finalS, Q, rot_list) from a text file, runs getOriTraj from T = inittime * ones(N) and prints opt_x, T,
// the final cost and the iteration count.  Host only: tests/test_capi_host.py compiles and runs it WITHOUT a GPU.
#include <array>
#include <cstdio>
#include <vector>

#include "svsdf.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = std::fopen(argv[1], "r");
    if (!f) return 3;
    int N;
    double inittime;
    if (std::fscanf(f, "%d %lf", &N, &inittime) != 2) return 4;
    double initS[9], finalS[9];
    for (double &v : initS) if (std::fscanf(f, "%lf", &v) != 1) return 4;
    for (double &v : finalS) if (std::fscanf(f, "%lf", &v) != 1) return 4;
    std::vector<std::array<double, 3>> Q(N - 1);
    std::vector<std::array<double, 9>> rot(N - 1);
    for (auto &q : Q) for (double &v : q) if (std::fscanf(f, "%lf", &v) != 1) return 4;
    for (auto &r : rot) for (double &v : r) if (std::fscanf(f, "%lf", &v) != 1) return 4;
    std::fclose(f);
    svsdf::OriTraj ori;  // conf = config/star.yaml's values
    std::vector<double> T(N, inittime), traj_T, traj_coeffs, opt_x;
    const bool ok = ori.getOriTraj(initS, finalS, Q, T, rot, N, traj_T, traj_coeffs, opt_x);
    std::printf("%d %.17g %d\n", ok ? 1 : 0, ori.final_cost, ori.iter);
    for (double v : opt_x) std::printf("%.17g ", v);
    std::printf("\n");
    for (double v : traj_T) std::printf("%.17g ", v);
    std::printf("\n");
    return 0;
}
