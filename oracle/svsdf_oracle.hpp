// oracle/svsdf_oracle.hpp — TEST INFRASTRUCTURE ONLY.
//
// CPU (double, optional OpenMP) restatement of the reference's SVSDF collision cost + gradient path.
// It is the checker for the CUDA product path and the timed "cpu_baseline"/"--impl reference" leg of
// bench.py.  Nothing under implicit_svsdf_planner_b200/ may include, link or call this file.
//
// PARITY PINNED TO THE REFERENCE'S OWN SOURCE (round 2): the reference (ZJU-FAST-Lab/Implicit-SVSDF-Planner @ f18fd91)
// ships no golden vectors for this path and its headers need Eigen/ROS/PCL (absent here), but the arithmetic itself
// compiles: `make -C oracle ref_path` builds oracle/_ref/libref_path_*.so from the reference tree where it lies
// (trajectory.hpp, minco.hpp whole; the Shape.hpp classes, the SweptVolumeManager query methods and the TrajOptimizer
// penalty loop cut verbatim by oracle/ref_extract.py; Eigen = the stand-in in oracle/ref_shim).  tests/test_oracle_ref_pin.py
// holds this restatement to it: every per-point output (Piece<5> samples, sdf, t*, gradient, outside and GSIP points,
// smoothedL1, tau<->T) BIT FOR BIT, sums (cost, gradC, gradT, f, g, MINCO) to summation-order rounding; fixtures of the
// reference's outputs are committed as tests/golden/ref_pin_*.npz.  Also: a trace of the reference's own LMBM binary driving
// this code (tests/golden/lmbm_trace_star_400.npz) and, for the mesh functor, the reference's fast-winding-number code
// compiled into oracle/_ref (tests/golden/fwn_ref.npz).
//
// Every function cites the reference file:line it follows (paths relative to /root/reference/src).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "portable_sincos.hpp"
#include "shapes.hpp"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Piece<5> / Trajectory<5>: utils/include/utils/trajectory.hpp
// ---------------------------------------------------------------------------------------------
struct Piece {
    double dur;
    double c[3][6];  // coeffMat 3x6, column 0 = t^5 ... column 5 = t^0 (trajectory.hpp:41,47)

    // trajectory.hpp:104-114 (ascending powers, tn *= t; not Horner)
    void getPos(double t, double p[3]) const {
        p[0] = p[1] = p[2] = 0.0;
        double tn = 1.0;
        for (int i = 5; i >= 0; i--) {
            for (int d = 0; d < 3; ++d) p[d] += tn * c[d][i];
            tn *= t;
        }
    }
    // trajectory.hpp:116-128
    void getVel(double t, double v[3]) const {
        v[0] = v[1] = v[2] = 0.0;
        double tn = 1.0;
        int n = 1;
        for (int i = 4; i >= 0; i--) {
            for (int d = 0; d < 3; ++d) v[d] += n * tn * c[d][i];
            tn *= t;
            n++;
        }
    }
};

struct Trajectory {
    std::vector<Piece> pieces;
    int getPieceNum() const { return (int)pieces.size(); }
    // trajectory.hpp:410-419
    double getTotalDuration() const {
        double total = 0.0;
        for (size_t i = 0; i < pieces.size(); ++i) total += pieces[i].dur;
        return total;
    }
    // trajectory.hpp:498-516 (t becomes piece-local; strict '>')
    int locatePieceIdx(double &t) const {
        int N = getPieceNum();
        int idx;
        double dur;
        for (idx = 0; idx < N && t > (dur = pieces[idx].dur); idx++) t -= dur;
        if (idx == N) {
            idx--;
            t += pieces[idx].dur;
        }
        return idx;
    }
    void getPos(double t, double p[3]) const {  // :518-522
        int i = locatePieceIdx(t);
        pieces[i].getPos(t, p);
    }
    void getVel(double t, double v[3]) const {  // :524-528
        int i = locatePieceIdx(t);
        pieces[i].getVel(t, v);
    }
};

// MINCO b (6N x 3, column-major: element (6i+k, d) at d*6N + 6i + k) -> Trajectory
// minco.hpp:515-528: piece i coeffMat = b.block<6,3>(6i,0).transpose().rowwise().reverse()
inline void trajectory_from_coeffs(int N, const double *T, const double *b, Trajectory &traj) {
    traj.pieces.resize(N);
    for (int i = 0; i < N; ++i) {
        traj.pieces[i].dur = T[i];
        for (int d = 0; d < 3; ++d)
            for (int k = 0; k < 6; ++k) traj.pieces[i].c[d][5 - k] = b[d * 6 * N + 6 * i + k];
    }
}

// ---------------------------------------------------------------------------------------------
// SweptVolumeManager: swept_volume/include/swept_volume/sw_manager.hpp
// ---------------------------------------------------------------------------------------------
struct SweptVolume {
    Shape shape;
    Trajectory traj;
    double traj_duration = 0.0;
    mutable uint64_t eval_count = 0;  // number of getSDFAtTimeStamp evaluations (per thread-unsafe; use
                                      // only single threaded when counting)
    bool count_evals = false;

    // sw_manager.hpp:376-385
    void updateTraj(const Trajectory &t) {
        traj = t;
        double td = traj.getTotalDuration();
        if (td < 3 * 1e2) traj_duration = td;
    }

    // getStateOnTrajStamp :465-474 + posEva2Rel :521-526 + getonlySDF — i.e. getSDFAtTimeStamp<false>
    // :741-757.  R = AngleAxis(yaw, Z) = [[c,-s,0],[s,c,0],[0,0,1]]; rel = R^T (p - xt).
    void relAt(const double p[3], double t, double rel[3]) const {
        double xt[3];
        traj.getPos(t, xt);
        double yaw = xt[2];
        double s, c;
        psc::sincos(yaw, s, c);
        // the mesh functor goes through getSDFAtTimeStamp_igl (:760-777), which zeroes the pose's third component (the
        // yaw) before posEva2Rel (`xt(2)=0`, :767); the analytic path leaves it in (its functors ignore z)
        if (shape.id == SH_MESH) xt[2] = 0;
        double d0 = p[0] - xt[0], d1 = p[1] - xt[1], d2 = p[2] - xt[2];
        rel[0] = c * d0 + s * d1 + 0.0 * d2;
        rel[1] = -s * d0 + c * d1 + 0.0 * d2;
        rel[2] = 0.0 * d0 + 0.0 * d1 + 1.0 * d2;
    }
    double sdfAt(const double p[3], double t) const {
        if (count_evals) {
#pragma omp atomic
            eval_count++;
        }
        double rel[3];
        relAt(p, t, rel);
        return shape_sdf(shape, rel[0], rel[1], rel[2]);
    }
    // getSDF_DOTAtTimeStamp :798-806 (central FD; analytic code below the early return is dead)
    double sdfDotAt(const double p[3], double t) const {
        double t1 = std::max(0.0, t - 0.000001);
        double t2 = std::min(traj_duration, t + 0.000001);
        double sdf1 = sdfAt(p, t1);
        double sdf2 = sdfAt(p, t2);
        return (sdf2 - sdf1) * 500000;
    }
    // getGradPrelAtTimeStamp :779-795
    void gradPrelAt(const double p[3], double t, double g[3]) const {
        if (count_evals) {
#pragma omp atomic
            eval_count += 4;
        }
        double rel[3];
        relAt(p, t, rel);
        shape_grad1(shape, rel[0], rel[1], rel[2], g);
    }

    // choiceTInit<false>(p, dt) :538-581
    double choiceTInit(const double p[3], double dt) const {
        double min_dis = 1e9, dis = 1e9, time_seed = 0.0;
        int pricision_layers = 4, current_layer = 1;
        double loop_terminal = traj_duration;
        double t = 0.0;
        while (current_layer <= pricision_layers) {
            if (current_layer == 1) t = 0.0;
            if (current_layer > 1) {
                t = std::max(0.0, time_seed - 10 * dt);
                loop_terminal = std::min(traj_duration, time_seed + 10 * dt);
            }
            for (; t <= loop_terminal; t += dt) {
                dis = sdfAt(p, t);
                if (dis < min_dis) {
                    time_seed = t;
                    min_dis = dis;
                }
            }
            dt *= 0.1;
            current_layer += 1;
        }
        return time_seed;
    }

    // gradientDescent :1249-1325 (momentum unused)
    void gradientDescent(double t_min, double t_max, const double x0, double &fx, double &x,
                         const double p[3]) const {
        int max_iter = 1000;
        double alpha = 0.01, tau = alpha, g = 0.0, tol = 1e-16;
        x = x0;
        double change = 0;
        double prev_x = 10000000.0;
        int iter = 0;
        bool stop = false;
        double x_candidate, fx_candidate;
        g = 100.0;
        while (iter < max_iter && !stop && std::abs(x - prev_x) > tol) {
            if (iter == 0) fx = sdfAt(p, x);
            g = sdfDotAt(p, x);
            tau = alpha;
            prev_x = x;
            for (int div = 1; div < 30; div++) {
                iter = iter + 1;
                g = sdfDotAt(p, x);
                change = -tau * ((int)(g > 0) - (g < 0));
                x_candidate = x + change;
                x_candidate = std::max(std::min(x_candidate, t_max), t_min);
                fx_candidate = sdfAt(p, x_candidate);
                if ((fx_candidate - fx) < 0) {
                    x = x_candidate;
                    fx = fx_candidate;
                    break;
                }
                tau = 0.5 * tau;
                if (div == 29) stop = true;
            }
        }
    }

    // getSDFofSweptVolume<false,true> :844-866
    double getSDFofSweptVolume(const double p[3], double &time_seed_f, double grad_prel[3],
                               bool need_grad = true) const {
        double ts, t_star = 0.0, sdf_star = 0.0;
        double dtime = 0.15;
        ts = choiceTInit(p, dtime);
        double tmin_ = std::max(0.0, ts - 3.4);
        double tmax_ = std::min(ts + 3.4, traj_duration);
        gradientDescent(tmin_, tmax_, ts, sdf_star, t_star, p);
        if (need_grad) gradPrelAt(p, t_star, grad_prel);
        time_seed_f = t_star;
        return sdf_star;
    }

    // getTrueSDFofSweptVolume<true> :916-1018 with SampleSet2D :41-124 / CircleCoord2D :25-40
    double getTrueSDFofSweptVolume(const double p[3], double &time_seed_f, double grad_prel[3],
                                   int *gsip_rounds = nullptr) const {
        const double PI = 3.14159265358979323846;  // Shape.hpp:31 (the 'PI' used by SampleSet2D)
        if (gsip_rounds) *gsip_rounds = 0;
        double argmin_dis = getSDFofSweptVolume(p, time_seed_f, grad_prel);
        if (argmin_dis > 0) return argmin_dis;

        double r0 = 10;
        double vel[3];
        traj.getVel(time_seed_f, vel);
        auto norm3 = [](const double v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
        if (norm3(vel) < 0.01) {  // :929-954
            if (time_seed_f < 0.1) {
                for (double t_scan = time_seed_f; t_scan <= traj_duration; t_scan += 0.1) {
                    traj.getVel(t_scan, vel);
                    if (norm3(vel) >= 0.01) break;
                }
            } else if (time_seed_f > traj_duration - 0.1) {
                for (double t_scan = time_seed_f; t_scan >= 0; t_scan -= 0.1) {
                    traj.getVel(t_scan, vel);
                    if (norm3(vel) >= 0.01) break;
                }
            }
        }
        // SampleSet2D::initSet :74-103
        double cx = p[0], cy = p[1];
        double r = r0;
        double theta0 = psc::atan2(vel[0], -vel[1]);
        if (theta0 < 0) theta0 += 2 * PI;
        double theta_res = PI + 0.1;
        const double rk_res = 1.5, rk0 = 1.0;

        double r_star = 0, max_g, cur_g, real_t_star = 0;
        double star_rk = 0, star_theta = 0;
        int iter = 1;
        double yk3[3], gtmp[3];
        while (true) {
            max_g = -100000;
            // getElements :59-71
            for (double rk = rk0; rk > 0; rk -= rk_res) {
                for (double theta = theta0; theta < theta0 + 2 * PI; theta += theta_res) {
                    // CircleCoord2D::getPosition :36-39
                    double sth, cth;
                    psc::sincos(theta, sth, cth);
                    yk3[0] = cx + rk * r * cth;
                    yk3[1] = cy + rk * r * sth;
                    yk3[2] = 0.0;
                    cur_g = getSDFofSweptVolume(yk3, time_seed_f, gtmp);
                    if (cur_g > max_g) {
                        max_g = cur_g;
                        real_t_star = time_seed_f;
                        star_rk = rk;
                        star_theta = theta;
                    }
                }
            }
            r_star = r - max_g;
            r = r_star;
            if (gsip_rounds) (*gsip_rounds)++;
            if (iter > 8) break;
            if (std::abs(max_g) < 0.1) break;
            // expandSet(2, yk_star.theta) :105-123
            theta_res /= (2 + 1);
            theta_res = std::max(0.3, theta_res);
            theta0 = star_theta;
            iter++;
        }
        double sst, cst;
        psc::sincos(star_theta, sst, cst);
        double corx = cx + star_rk * r_star * cst;
        double cory = cy + star_rk * r_star * sst;
        double gx = corx - p[0], gy = cory - p[1], gz = 0.0;
        double z = gx * gx + gy * gy + gz * gz;
        if (z > 0) {  // Eigen normalize()
            double n = std::sqrt(z);
            gx /= n; gy /= n; gz /= n;
        }
        grad_prel[0] = gx; grad_prel[1] = gy; grad_prel[2] = gz;
        time_seed_f = real_t_star;
        return -r_star;
    }
};

// ---------------------------------------------------------------------------------------------
// TrajOptimizer pieces: planner_algorithm/include/planner_algorithm/back_end_optimizer.hpp
// ---------------------------------------------------------------------------------------------
// smoothedL1 :316-340
inline bool smoothedL1(const double x, const double mu, double &f, double &df) {
    if (x < 0.0) return false;
    else if (x > mu) {
        f = x - 0.5 * mu;
        df = 1.0;
        return true;
    } else {
        const double xdmu = x / mu;
        const double sqrxdmu = xdmu * xdmu;
        const double mumxd2 = mu - 0.5 * x;
        f = mumxd2 * sqrxdmu * xdmu;
        df = sqrxdmu * ((-0.5) * xdmu + 3.0 * mumxd2 / mu);
        return true;
    }
}

struct PointResult {  // per query point record (golden-vector format)
    double sdf, tstar, g[3];  // g: body-frame gradient as used by grad_cost_p_sw (after the :832 rotation)
    int piece;
    double pena;  // weight_p * L
};

struct CostParams {
    double weight_p = 60.0;
    double safety_hor = 0.7;
    int threads = 1;
};

// addSaftyPenaOnSweptVolumeParallelTrueSDF :774-869 (+ grad_cost_p_sw :1031-1066)
// T: N durations; coeffs: MINCO b 6N x 3 column-major; accumulates into cost / gradT[N] / gradC[6N x 3].
inline void addSafetyPenaltyTrueSDF(const SweptVolume &sv, const CostParams &cp, int N, const double *T,
                                    const double *coeffs, const double *points, int64_t P, int stride,
                                    double &cost, double *gradT, double *gradC, PointResult *per_point,
                                    int64_t *n_inside = nullptr) {
    const double weightPos = cp.weight_p;
    int64_t inside = 0;
    (void)T;
#pragma omp parallel for num_threads(cp.threads) schedule(dynamic) reduction(+ : inside)
    for (int64_t k = 0; k < P; ++k) {
        double pos_eva[3] = {points[k * stride], points[k * stride + 1], 0.0};  // :790-791
        double gradp_rel[3] = {0, 0, 0};
        double time_star = 0.0;
        double sdf_value = sv.getTrueSDFofSweptVolume(pos_eva, time_star, gradp_rel);
        if (!(sdf_value > 0)) inside++;
        double time_local = time_star;
        int i = sv.traj.locatePieceIdx(time_local);  // :797
        double s1 = time_local, s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
        double beta0[6] = {1.0, s1, s2, s3, s4, s5};
        double beta1[6] = {0.0, 1.0, 2.0 * s1, 3.0 * s2, 4.0 * s3, 5.0 * s4};
        double pos[3], vel[3];
        for (int d = 0; d < 3; ++d) {  // c^T beta : c = coeffs.block<6,3>(6i,0)
            const double *c = coeffs + d * 6 * N + 6 * i;
            double a = 0, b = 0;
            for (int q = 0; q < 6; ++q) { a += c[q] * beta0[q]; b += c[q] * beta1[q]; }
            pos[d] = a; vel[d] = b;
        }
        double yaw = pos[2];
        double sy, cy;
        psc::sincos(yaw, sy, cy);
        // rotate = [[cy,-sy,0],[sy,cy,0],[0,0,1]]
        pos[2] = 0.0;  // :829
        if (sdf_value < 0) {  // :832 gradp_rel = rotate^T * gradp_rel
            double g0 = cy * gradp_rel[0] + sy * gradp_rel[1];
            double g1 = -sy * gradp_rel[0] + cy * gradp_rel[1];
            gradp_rel[0] = g0; gradp_rel[1] = g1;
        }
        // grad_cost_p_sw :1031-1066
        double costp = 0.0, gradp[3] = {0, 0, 0}, grad_yaw_v = 0.0;
        double sdf_cost = -1.0, sdf_out_grad = 0.0;
        smoothedL1(cp.safety_hor - sdf_value, 0.01, sdf_cost, sdf_out_grad);
        // sdf_grad = -L' * ( -(I) * rotate * g )
        double rg0 = -(cy * gradp_rel[0] + (-sy) * gradp_rel[1]);
        double rg1 = -(sy * gradp_rel[0] + cy * gradp_rel[1]);
        double sdf_grad0 = -sdf_out_grad * rg0, sdf_grad1 = -sdf_out_grad * rg1;
        bool active = false;
        if (sdf_cost > 0) {
            costp += sdf_cost;
            gradp[0] += sdf_grad0;
            gradp[1] += sdf_grad1;
            double d0 = pos_eva[0] - pos[0], d1 = pos_eva[1] - pos[1];
            // VR_theta^T * (p - x): VR=[[-s,-c,0],[c,-s,0],[0,0,1]]
            double w0 = -sy * d0 + cy * d1;
            double w1 = -cy * d0 + -sy * d1;
            grad_yaw_v = (-sdf_out_grad * gradp_rel[0]) * w0 + (-sdf_out_grad * gradp_rel[1]) * w1;
            active = costp > 0;
        }
        double gradPos[3] = {0, 0, 0}, grad_yaw = 0.0, pena = 0.0;
        if (active) {
            gradPos[0] += weightPos * gradp[0];
            gradPos[1] += weightPos * gradp[1];
            grad_yaw += weightPos * grad_yaw_v;
            pena += weightPos * costp;
        }
        double G[3] = {gradPos[0], gradPos[1], grad_yaw};
        double gdT = -(G[0] * vel[0] + G[1] * vel[1] + G[2] * vel[2]);
        if (per_point) {
            PointResult &pr = per_point[k];
            pr.sdf = sdf_value; pr.tstar = time_star;
            pr.g[0] = gradp_rel[0]; pr.g[1] = gradp_rel[1]; pr.g[2] = gradp_rel[2];
            pr.piece = i; pr.pena = pena;
        }
#pragma omp critical
        {
            cost += pena;
            for (int d = 0; d < 3; ++d)
                for (int q = 0; q < 6; ++q) gradC[d * 6 * N + 6 * i + q] += beta0[q] * G[d];
            for (int j = 0; j < i; ++j) gradT[j] += gdT;  // :859-862 (j < i only)
        }
    }
    if (n_inside) *n_inside = inside;
}

}  // namespace oracle
