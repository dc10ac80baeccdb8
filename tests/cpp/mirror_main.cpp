// tests/cpp/mirror_main.cpp — drives the C++ mirror (include/svsdf.hpp) the way the reference's plan_manager drives
// TrajOptimizer / SweptVolumeManager (plan_manager.cpp:47-94, 128-227) and prints results for tests/test_gpu_cpp_mirror.py.
// Input file (text): shape N P weight_p safety_hor rho | T[N] | coeffs[18N] (col-major) | initS[9] finalS[9] (col-major) |
// x0[N+3(N-1)] | points P x 3
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <vector>

#include "svsdf.hpp"

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: mirror_main <input.txt>\n"); return 2; }
    std::ifstream in(argv[1]);
    std::string shape;
    int N, P;
    svsdf::Config config;
    in >> shape >> N >> P >> config.weight_p >> config.safety_hor >> config.rho;
    config.inputdata = "shapes/" + shape + ".obj";  // yaml `inputdata`
    std::vector<double> T(N), coeffs(18 * N), initS(9), finalS(9), x0(N + 3 * (N - 1));
    for (auto &v : T) in >> v;
    for (auto &v : coeffs) in >> v;
    for (auto &v : initS) in >> v;
    for (auto &v : finalS) in >> v;
    for (auto &v : x0) in >> v;

    // PlannerManager::init (plan_manager.cpp:47-94)
    auto sv_manager = std::make_shared<svsdf::SweptVolumeManager>(config);
    auto minco_traj_optimizer = std::make_shared<svsdf::TrajOptimizer>();
    minco_traj_optimizer->setParam(config);
    minco_traj_optimizer->setEnvironment(sv_manager);
    // generateTraj (plan_manager.cpp:168-175)
    minco_traj_optimizer->parallel_points.resize(P);
    for (int i = 0; i < P; ++i) in >> minco_traj_optimizer->parallel_points[i][0] >> minco_traj_optimizer->parallel_points[i][1] >> minco_traj_optimizer->parallel_points[i][2];
    minco_traj_optimizer->parallel_points_num = P;
    if (minco_traj_optimizer->uploadPoints() != SVSDF_OK) { std::fprintf(stderr, "%s\n", sv_manager->last_error()); return 1; }

    std::printf("%.17g\n", (double)svsdf::shape::registry_id(shape));
    // shape functor
    const double rel[3] = {0.0, 3.8, 0.0};
    double g1[3];
    std::printf("%.17g\n", sv_manager->current_robot_shape->getSDFwithGrad1(rel, g1));
    std::printf("%.17g %.17g %.17g\n", g1[0], g1[1], g1[2]);
    // per-point query on the first 5 points
    sv_manager->updateTraj(N, T.data(), coeffs.data());
    for (int i = 0; i < 5 && i < P; ++i) {
        double pos_eva[3] = {minco_traj_optimizer->parallel_points[i][0], minco_traj_optimizer->parallel_points[i][1], 0.0};
        double time_seed_f = 0.0, grad_prel[3];
        double sdf = sv_manager->getTrueSDFofSweptVolume(pos_eva, time_seed_f, grad_prel, false);
        std::printf("%.17g %.17g %.17g %.17g\n", sdf, time_seed_f, grad_prel[0], grad_prel[1]);
    }
    // the OpenMP loop's replacement (accumulating)
    double cost = 1.5;
    std::vector<double> gradT(N, 0.25), gradC(18 * N, -0.5);
    int rc = svsdf::TrajOptimizer::addSaftyPenaOnSweptVolumeParallelTrueSDF(minco_traj_optimizer.get(), N, T.data(), coeffs.data(), cost, gradT.data(), gradC.data());
    std::printf("%d %.17g\n", rc, cost);
    for (double v : gradT) std::printf("%.17g ", v);
    std::printf("\n");
    for (double v : gradC) std::printf("%.17g ", v);
    std::printf("\n");
    // the LMBM-compatible callback
    minco_traj_optimizer->setConditions(initS.data(), finalS.data(), N);
    std::vector<double> g(x0.size());
    double f = svsdf::TrajOptimizer::costFunctionLmbmParallel(minco_traj_optimizer.get(), x0.data(), g.data(), (int)x0.size());
    std::printf("%.17g %.17g %.17g %.17g\n", f, minco_traj_optimizer->cost_pos, minco_traj_optimizer->cost_other, minco_traj_optimizer->cost_total);
    for (double v : g) std::printf("%.17g ", v);
    std::printf("\n");
    // optimize_traj_lmbm
    std::vector<double> opt_x = x0, trajT, trajC;
    svsdf_opt_stats st;
    config.past = 3; config.relCostTol = 1e-5;
    minco_traj_optimizer->setParam(config);
    int ret = minco_traj_optimizer->optimize_traj_lmbm(initS.data(), finalS.data(), opt_x, N, trajT, trajC, &st);
    std::printf("%d %d %d %.17g\n", ret, st.iterations, st.evaluations, st.final_cost);
    return 0;
}
