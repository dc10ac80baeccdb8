// oracle/minco_oracle.hpp — TEST INFRASTRUCTURE ONLY (see svsdf_oracle.hpp header).
// MINCO_S3NU, the tau<->T maps and the cost callback are pinned to the reference's own minco.hpp / back_end_optimizer.hpp
// code (oracle/_ref/libref_path_*.so, tests/test_oracle_ref_pin.py: b, energy, gradients, adjoint to 1e-13 relative —
// Eigen's reduction order inside getEnergy/propogateGrad is the only freedom); the L-BFGS is a restatement of lbfgs_ref.hpp..
//
// Restatement of  utils/include/utils/minco.hpp (BandedSystem :43-198, MINCO_S3NU :397-655),
// the tau<->T maps and cost wrapper of planner_algorithm/include/planner_algorithm/back_end_optimizer.hpp
// (:163-314, :344-408) and the clean L-BFGS of utils/include/utils/lbfgs_ref.hpp (:276-716).
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

#include "svsdf_oracle.hpp"

namespace oracle {

// minco.hpp:43-198 — banded LU without pivoting, storage ptrData[(i-j+upperBw)*N + j]
class BandedSystem {
   public:
    void create(int n, int p, int q) {
        N = n; lowerBw = p; upperBw = q;
        data.assign((size_t)N * (lowerBw + upperBw + 1), 0.0);
    }
    void reset() { std::fill(data.begin(), data.end(), 0.0); }
    double &operator()(int i, int j) { return data[(size_t)(i - j + upperBw) * N + j]; }
    const double &operator()(int i, int j) const { return data[(size_t)(i - j + upperBw) * N + j]; }
    void factorizeLU() {  // :99-131
        int iM, jM;
        double cVl;
        for (int k = 0; k <= N - 2; k++) {
            iM = std::min(k + lowerBw, N - 1);
            cVl = (*this)(k, k);
            for (int i = k + 1; i <= iM; i++)
                if ((*this)(i, k) != 0.0) (*this)(i, k) /= cVl;
            jM = std::min(k + upperBw, N - 1);
            for (int j = k + 1; j <= jM; j++) {
                cVl = (*this)(k, j);
                if (cVl != 0.0)
                    for (int i = k + 1; i <= iM; i++)
                        if ((*this)(i, k) != 0.0) (*this)(i, j) -= (*this)(i, k) * cVl;
            }
        }
    }
    // b: N x m, column-major (b[c*N + r])
    void solve(double *b, int m) const {  // :136-164
        int iM;
        for (int j = 0; j <= N - 1; j++) {
            iM = std::min(j + lowerBw, N - 1);
            for (int i = j + 1; i <= iM; i++)
                if ((*this)(i, j) != 0.0)
                    for (int c = 0; c < m; ++c) b[c * N + i] -= (*this)(i, j) * b[c * N + j];
        }
        for (int j = N - 1; j >= 0; j--) {
            for (int c = 0; c < m; ++c) b[c * N + j] /= (*this)(j, j);
            iM = std::max(0, j - upperBw);
            for (int i = iM; i <= j - 1; i++)
                if ((*this)(i, j) != 0.0)
                    for (int c = 0; c < m; ++c) b[c * N + i] -= (*this)(i, j) * b[c * N + j];
        }
    }
    void solveAdj(double *b, int m) const {  // :169-197
        int iM;
        for (int j = 0; j <= N - 1; j++) {
            for (int c = 0; c < m; ++c) b[c * N + j] /= (*this)(j, j);
            iM = std::min(j + upperBw, N - 1);
            for (int i = j + 1; i <= iM; i++)
                if ((*this)(j, i) != 0.0)
                    for (int c = 0; c < m; ++c) b[c * N + i] -= (*this)(j, i) * b[c * N + j];
        }
        for (int j = N - 1; j >= 0; j--) {
            iM = std::max(0, j - lowerBw);
            for (int i = iM; i <= j - 1; i++)
                if ((*this)(j, i) != 0.0)
                    for (int c = 0; c < m; ++c) b[c * N + i] -= (*this)(j, i) * b[c * N + j];
        }
    }

   private:
    int N = 0, lowerBw = 0, upperBw = 0;
    std::vector<double> data;
};

// minco.hpp:397-655
class MincoS3NU {
   public:
    int N = 0;
    double head[3][3], tail[3][3];  // [dim][derivative]: headstate.col(k) = k-th derivative
    std::vector<double> b;           // 6N x 3 column-major
    std::vector<double> T1, T2, T3, T4, T5;
    BandedSystem A;

    double &B(int r, int d) { return b[(size_t)d * 6 * N + r]; }
    const double &B(int r, int d) const { return b[(size_t)d * 6 * N + r]; }

    // setConditions :413-431. headState/tailState: 3x3 column-major (col k = k-th derivative)
    void setConditions(const double *headState, const double *tailState, int pieceNum) {
        N = pieceNum;
        for (int k = 0; k < 3; ++k)
            for (int d = 0; d < 3; ++d) {
                head[d][k] = headState[k * 3 + d];
                tail[d][k] = tailState[k * 3 + d];
            }
        A.create(6 * N, 6, 6);
        b.assign((size_t)6 * N * 3, 0.0);
        T1.resize(N); T2.resize(N); T3.resize(N); T4.resize(N); T5.resize(N);
    }
    // setParameters :433-513. inPs: 3 x (N-1) column-major
    void setParameters(const double *inPs, const double *ts) {
        for (int i = 0; i < N; ++i) {
            T1[i] = ts[i];
            T2[i] = T1[i] * T1[i];
            T3[i] = T2[i] * T1[i];
            T4[i] = T2[i] * T2[i];
            T5[i] = T4[i] * T1[i];
        }
        A.reset();
        std::fill(b.begin(), b.end(), 0.0);
        A(0, 0) = 1.0; A(1, 1) = 1.0; A(2, 2) = 2.0;
        for (int d = 0; d < 3; ++d) { B(0, d) = head[d][0]; B(1, d) = head[d][1]; B(2, d) = head[d][2]; }
        for (int i = 0; i < N - 1; i++) {
            A(6 * i + 3, 6 * i + 3) = 6.0;
            A(6 * i + 3, 6 * i + 4) = 24.0 * T1[i];
            A(6 * i + 3, 6 * i + 5) = 60.0 * T2[i];
            A(6 * i + 3, 6 * i + 9) = -6.0;
            A(6 * i + 4, 6 * i + 4) = 24.0;
            A(6 * i + 4, 6 * i + 5) = 120.0 * T1[i];
            A(6 * i + 4, 6 * i + 10) = -24.0;
            A(6 * i + 5, 6 * i) = 1.0;
            A(6 * i + 5, 6 * i + 1) = T1[i];
            A(6 * i + 5, 6 * i + 2) = T2[i];
            A(6 * i + 5, 6 * i + 3) = T3[i];
            A(6 * i + 5, 6 * i + 4) = T4[i];
            A(6 * i + 5, 6 * i + 5) = T5[i];
            A(6 * i + 6, 6 * i) = 1.0;
            A(6 * i + 6, 6 * i + 1) = T1[i];
            A(6 * i + 6, 6 * i + 2) = T2[i];
            A(6 * i + 6, 6 * i + 3) = T3[i];
            A(6 * i + 6, 6 * i + 4) = T4[i];
            A(6 * i + 6, 6 * i + 5) = T5[i];
            A(6 * i + 6, 6 * i + 6) = -1.0;
            A(6 * i + 7, 6 * i + 1) = 1.0;
            A(6 * i + 7, 6 * i + 2) = 2 * T1[i];
            A(6 * i + 7, 6 * i + 3) = 3 * T2[i];
            A(6 * i + 7, 6 * i + 4) = 4 * T3[i];
            A(6 * i + 7, 6 * i + 5) = 5 * T4[i];
            A(6 * i + 7, 6 * i + 7) = -1.0;
            A(6 * i + 8, 6 * i + 2) = 2.0;
            A(6 * i + 8, 6 * i + 3) = 6 * T1[i];
            A(6 * i + 8, 6 * i + 4) = 12 * T2[i];
            A(6 * i + 8, 6 * i + 5) = 20 * T3[i];
            A(6 * i + 8, 6 * i + 8) = -2.0;
            for (int d = 0; d < 3; ++d) B(6 * i + 5, d) = inPs[i * 3 + d];
        }
        A(6 * N - 3, 6 * N - 6) = 1.0;
        A(6 * N - 3, 6 * N - 5) = T1[N - 1];
        A(6 * N - 3, 6 * N - 4) = T2[N - 1];
        A(6 * N - 3, 6 * N - 3) = T3[N - 1];
        A(6 * N - 3, 6 * N - 2) = T4[N - 1];
        A(6 * N - 3, 6 * N - 1) = T5[N - 1];
        A(6 * N - 2, 6 * N - 5) = 1.0;
        A(6 * N - 2, 6 * N - 4) = 2 * T1[N - 1];
        A(6 * N - 2, 6 * N - 3) = 3 * T2[N - 1];
        A(6 * N - 2, 6 * N - 2) = 4 * T3[N - 1];
        A(6 * N - 2, 6 * N - 1) = 5 * T4[N - 1];
        A(6 * N - 1, 6 * N - 4) = 2;
        A(6 * N - 1, 6 * N - 3) = 6 * T1[N - 1];
        A(6 * N - 1, 6 * N - 2) = 12 * T2[N - 1];
        A(6 * N - 1, 6 * N - 1) = 20 * T3[N - 1];
        for (int d = 0; d < 3; ++d) {
            B(6 * N - 3, d) = tail[d][0];
            B(6 * N - 2, d) = tail[d][1];
            B(6 * N - 1, d) = tail[d][2];
        }
        A.factorizeLU();
        A.solve(b.data(), 3);
    }
    double rowdot(int r1, int r2) const {
        return B(r1, 0) * B(r2, 0) + B(r1, 1) * B(r2, 1) + B(r1, 2) * B(r2, 2);
    }
    // getEnergy :530-543
    double getEnergy() const {
        double energy = 0.0;
        for (int i = 0; i < N; i++) {
            energy += 36.0 * rowdot(6 * i + 3, 6 * i + 3) * T1[i] +
                      144.0 * rowdot(6 * i + 4, 6 * i + 3) * T2[i] +
                      192.0 * rowdot(6 * i + 4, 6 * i + 4) * T3[i] +
                      240.0 * rowdot(6 * i + 5, 6 * i + 3) * T3[i] +
                      720.0 * rowdot(6 * i + 5, 6 * i + 4) * T4[i] +
                      720.0 * rowdot(6 * i + 5, 6 * i + 5) * T5[i];
        }
        return energy;
    }
    // getEnergyPartialGradByCoeffs :550-567 (gdC 6N x 3 column-major)
    void getEnergyPartialGradByCoeffs(double *gdC) const {
        for (int i = 0; i < N; i++)
            for (int d = 0; d < 3; ++d) {
                double *g = gdC + (size_t)d * 6 * N + 6 * i;
                g[5] = 240.0 * B(6 * i + 3, d) * T3[i] + 720.0 * B(6 * i + 4, d) * T4[i] +
                       1440.0 * B(6 * i + 5, d) * T5[i];
                g[4] = 144.0 * B(6 * i + 3, d) * T2[i] + 384.0 * B(6 * i + 4, d) * T3[i] +
                       720.0 * B(6 * i + 5, d) * T4[i];
                g[3] = 72.0 * B(6 * i + 3, d) * T1[i] + 144.0 * B(6 * i + 4, d) * T2[i] +
                       240.0 * B(6 * i + 5, d) * T3[i];
                g[0] = g[1] = g[2] = 0.0;
            }
    }
    // getEnergyPartialGradByTimes :569-582
    void getEnergyPartialGradByTimes(double *gdT) const {
        for (int i = 0; i < N; i++) {
            gdT[i] = 36.0 * rowdot(6 * i + 3, 6 * i + 3) + 288.0 * rowdot(6 * i + 4, 6 * i + 3) * T1[i] +
                     576.0 * rowdot(6 * i + 4, 6 * i + 4) * T2[i] +
                     720.0 * rowdot(6 * i + 5, 6 * i + 3) * T2[i] +
                     2880.0 * rowdot(6 * i + 5, 6 * i + 4) * T3[i] +
                     3600.0 * rowdot(6 * i + 5, 6 * i + 5) * T4[i];
        }
    }
    // propogateGrad :584-654. gradByPoints 3 x (N-1) column-major, gradByTimes N
    void propogateGrad(const double *partialGradByCoeffs, const double *partialGradByTimes,
                       double *gradByPoints, double *gradByTimes) const {
        std::vector<double> adj(partialGradByCoeffs, partialGradByCoeffs + (size_t)6 * N * 3);
        A.solveAdj(adj.data(), 3);
        auto ADJ = [&](int r, int d) { return adj[(size_t)d * 6 * N + r]; };
        for (int i = 0; i < N - 1; i++)
            for (int d = 0; d < 3; ++d) gradByPoints[i * 3 + d] = ADJ(6 * i + 5, d);
        double B1[6][3];
        for (int i = 0; i < N - 1; i++) {
            for (int d = 0; d < 3; ++d) {
                B1[2][d] = -(B(i * 6 + 1, d) + 2.0 * T1[i] * B(i * 6 + 2, d) + 3.0 * T2[i] * B(i * 6 + 3, d) +
                             4.0 * T3[i] * B(i * 6 + 4, d) + 5.0 * T4[i] * B(i * 6 + 5, d));
                B1[3][d] = B1[2][d];
                B1[4][d] = -(2.0 * B(i * 6 + 2, d) + 6.0 * T1[i] * B(i * 6 + 3, d) +
                             12.0 * T2[i] * B(i * 6 + 4, d) + 20.0 * T3[i] * B(i * 6 + 5, d));
                B1[5][d] = -(6.0 * B(i * 6 + 3, d) + 24.0 * T1[i] * B(i * 6 + 4, d) + 60.0 * T2[i] * B(i * 6 + 5, d));
                B1[0][d] = -(24.0 * B(i * 6 + 4, d) + 120.0 * T1[i] * B(i * 6 + 5, d));
                B1[1][d] = -120.0 * B(i * 6 + 5, d);
            }
            double s = 0.0;  // Eigen .sum() of a 6x3 column-major expression: column by column
            for (int d = 0; d < 3; ++d)
                for (int r = 0; r < 6; ++r) s += B1[r][d] * ADJ(6 * i + 3 + r, d);
            gradByTimes[i] = s;
        }
        double B2[3][3];
        for (int d = 0; d < 3; ++d) {
            B2[0][d] = -(B(6 * N - 5, d) + 2.0 * T1[N - 1] * B(6 * N - 4, d) + 3.0 * T2[N - 1] * B(6 * N - 3, d) +
                         4.0 * T3[N - 1] * B(6 * N - 2, d) + 5.0 * T4[N - 1] * B(6 * N - 1, d));
            B2[1][d] = -(2.0 * B(6 * N - 4, d) + 6.0 * T1[N - 1] * B(6 * N - 3, d) +
                         12.0 * T2[N - 1] * B(6 * N - 2, d) + 20.0 * T3[N - 1] * B(6 * N - 1, d));
            B2[2][d] = -(6.0 * B(6 * N - 3, d) + 24.0 * T1[N - 1] * B(6 * N - 2, d) + 60.0 * T2[N - 1] * B(6 * N - 1, d));
        }
        double s = 0.0;
        for (int d = 0; d < 3; ++d)
            for (int r = 0; r < 3; ++r) s += B2[r][d] * ADJ(6 * N - 3 + r, d);
        gradByTimes[N - 1] = s;
        for (int i = 0; i < N; ++i) gradByTimes[i] += partialGradByTimes[i];
    }
};

// back_end_optimizer.hpp:213-226
inline double forwardT1(double tau) {
    return tau > 0.0 ? ((0.5 * tau + 1.0) * tau + 1.0) : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0);
}
// :228-241
inline double backwardT1(double T) {
    return T > 1.0 ? (std::sqrt(2.0 * T - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / T - 1.0));
}
// :268-289
inline double backwardGradT1(double tau, double gradT) {
    if (tau > 0) return gradT * (tau + 1.0);
    double denSqrt = (0.5 * tau - 1.0) * tau + 1.0;
    return gradT * (1.0 - tau) / (denSqrt * denSqrt);
}

// The TrajOptimizer state needed by costFunctionLmbmParallel :344-408
struct TrajOptimizerOracle {
    SweptVolume sv;
    CostParams cp;
    double rho = 3.8;
    int pieceN = 0;
    MincoS3NU minco;
    std::vector<double> points;  // query points, stride 3
    int64_t P = 0;
    // outputs of the last evaluation
    double cost_pos = 0, cost_other = 0, cost_total = 0;
    int64_t n_evals = 0;
    std::vector<double> times, gradByTimes, partialGradByTimes, partialGradByCoeffs, gradByPoints;

    void setConditions(const double *initS, const double *finalS, int N) {
        pieceN = N;
        minco.setConditions(initS, finalS, N);
        times.resize(N); gradByTimes.resize(N); partialGradByTimes.resize(N);
        partialGradByCoeffs.resize((size_t)18 * N);
        gradByPoints.resize((size_t)3 * (N - 1));
    }
    // costFunctionLmbmParallel :344-408.  x = [tau (N), xi (3(N-1))]
    double evaluate(const double *x, double *g) {
        const int N = pieceN;
        n_evals++;
        for (int i = 0; i < N; ++i) times[i] = forwardT1(x[i]);
        const double *q = x + N;  // forwardP :174-185 : P.col(i) = xi[3i..3i+2]
        minco.setParameters(q, times.data());
        double cost = minco.getEnergy();
        minco.getEnergyPartialGradByCoeffs(partialGradByCoeffs.data());
        minco.getEnergyPartialGradByTimes(partialGradByTimes.data());
        Trajectory tr;
        trajectory_from_coeffs(N, times.data(), minco.b.data(), tr);
        sv.updateTraj(tr);
        double energy_cost = cost;
        addSafetyPenaltyTrueSDF(sv, cp, N, times.data(), minco.b.data(), points.data(), P, 3, cost,
                                partialGradByTimes.data(), partialGradByCoeffs.data(), nullptr);
        double pos_cost = cost - energy_cost;
        minco.propogateGrad(partialGradByCoeffs.data(), partialGradByTimes.data(), gradByPoints.data(),
                            gradByTimes.data());
        double tsum = 0.0;
        for (int i = 0; i < N; ++i) tsum += times[i];
        cost += rho * tsum;
        for (int i = 0; i < N; ++i) gradByTimes[i] += rho;
        cost_pos = pos_cost; cost_other = cost - pos_cost; cost_total = cost;
        for (int i = 0; i < N; ++i) g[i] = backwardGradT1(x[i], gradByTimes[i]);
        for (int i = 0; i < 3 * (N - 1); ++i) g[N + i] = gradByPoints[i];
        return cost;
    }
};

// ---------------------------------------------------------------------------------------------
// lbfgs_ref.hpp (clean upstream LBFGS-Lite): line_search_lewisoverton :276-395, lbfgs_optimize :434-716
// ---------------------------------------------------------------------------------------------
struct LbfgsParams {
    int mem_size = 8;
    double g_epsilon = 1.0e-5;
    int past = 3;
    double delta = 1.0e-6;
    int max_iterations = 0;
    int max_linesearch = 64;
    double min_step = 1.0e-20;
    double max_step = 1.0e+20;
    double f_dec_coeff = 1.0e-4;
    double s_curv_coeff = 0.9;
    double cautious_factor = 1.0e-6;
    double machine_prec = 1.0e-16;
};
enum {
    LBFGS_CONVERGENCE = 0, LBFGS_STOP, LBFGS_CANCELED,
    LBFGSERR_UNKNOWNERROR = -1024, LBFGSERR_INVALID_N, LBFGSERR_INVALID_MEMSIZE, LBFGSERR_INVALID_GEPSILON,
    LBFGSERR_INVALID_TESTPERIOD, LBFGSERR_INVALID_DELTA, LBFGSERR_INVALID_MINSTEP, LBFGSERR_INVALID_MAXSTEP,
    LBFGSERR_INVALID_FDECCOEFF, LBFGSERR_INVALID_SCURVCOEFF, LBFGSERR_INVALID_MACHINEPREC,
    LBFGSERR_INVALID_MAXLINESEARCH, LBFGSERR_INVALID_FUNCVAL, LBFGSERR_MINIMUMSTEP, LBFGSERR_MAXIMUMSTEP,
    LBFGSERR_MAXIMUMLINESEARCH, LBFGSERR_MAXIMUMITERATION, LBFGSERR_WIDTHTOOSMALL,
    LBFGSERR_INVALIDPARAMETERS, LBFGSERR_INCREASEGRADIENT,
};
typedef std::function<double(const double *x, double *g)> EvalFn;
typedef std::function<int(const double *x, const double *g, double fx, double step, int k, int ls)> ProgressFn;

inline double vdot(const std::vector<double> &a, const std::vector<double> &b) {
    double s = 0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}

inline int line_search_lewisoverton(std::vector<double> &x, double &f, std::vector<double> &g, double &stp,
                                    const std::vector<double> &s, const std::vector<double> &xp,
                                    const std::vector<double> &gp, double stpmin, double stpmax,
                                    const EvalFn &eval, const LbfgsParams &param) {
    int count = 0;
    bool brackt = false, touched = false;
    double finit, dginit, dgtest, dstest;
    double mu = 0.0, nu = stpmax;
    if (!(stp > 0.0)) return LBFGSERR_INVALIDPARAMETERS;
    dginit = vdot(gp, s);
    if (0.0 < dginit) return LBFGSERR_INCREASEGRADIENT;
    finit = f;
    dgtest = param.f_dec_coeff * dginit;
    dstest = param.s_curv_coeff * dginit;
    const size_t n = x.size();
    while (true) {
        for (size_t i = 0; i < n; ++i) x[i] = xp[i] + stp * s[i];
        f = eval(x.data(), g.data());
        ++count;
        if (std::isinf(f) || std::isnan(f)) return LBFGSERR_INVALID_FUNCVAL;
        if (f > finit + stp * dgtest) {
            nu = stp;
            brackt = true;
        } else {
            if (vdot(g, s) < dstest) mu = stp;
            else return count;
        }
        if (param.max_linesearch <= count) return LBFGSERR_MAXIMUMLINESEARCH;
        if (brackt && (nu - mu) < param.machine_prec * nu) return LBFGSERR_WIDTHTOOSMALL;
        if (brackt) stp = 0.5 * (mu + nu);
        else stp *= 2.0;
        if (stp < stpmin) return LBFGSERR_MINIMUMSTEP;
        if (stp > stpmax) {
            if (touched) return LBFGSERR_MAXIMUMSTEP;
            touched = true;
            stp = stpmax;
        }
    }
}

inline int lbfgs_optimize(std::vector<double> &x, double &f, const EvalFn &eval, const ProgressFn &progress,
                          const LbfgsParams &param, int *iters_out = nullptr) {
    int ret, i, j, k, ls, end, bound;
    double step, step_min, step_max, fx, ys, yy;
    double gnorm_inf, xnorm_inf, beta, rate, cau;
    const int n = (int)x.size();
    const int m = param.mem_size;
    if (n <= 0) return LBFGSERR_INVALID_N;
    if (m <= 0) return LBFGSERR_INVALID_MEMSIZE;
    if (param.g_epsilon < 0.0) return LBFGSERR_INVALID_GEPSILON;
    if (param.past < 0) return LBFGSERR_INVALID_TESTPERIOD;
    if (param.delta < 0.0) return LBFGSERR_INVALID_DELTA;
    if (param.min_step < 0.0) return LBFGSERR_INVALID_MINSTEP;
    if (param.max_step < param.min_step) return LBFGSERR_INVALID_MAXSTEP;
    if (!(param.f_dec_coeff > 0.0 && param.f_dec_coeff < 1.0)) return LBFGSERR_INVALID_FDECCOEFF;
    if (!(param.s_curv_coeff < 1.0 && param.s_curv_coeff > param.f_dec_coeff)) return LBFGSERR_INVALID_SCURVCOEFF;
    if (!(param.machine_prec > 0.0)) return LBFGSERR_INVALID_MACHINEPREC;
    if (param.max_linesearch <= 0) return LBFGSERR_INVALID_MAXLINESEARCH;

    std::vector<double> xp(n), g(n), gp(n), d(n), pf(std::max(1, param.past));
    std::vector<double> lm_alpha(m, 0.0), lm_ys(m, 0.0);
    std::vector<std::vector<double>> lm_s(m, std::vector<double>(n, 0.0)), lm_y(m, std::vector<double>(n, 0.0));
    auto inf_norm = [](const std::vector<double> &v) { double r = 0; for (double e : v) r = std::max(r, std::fabs(e)); return r; };

    fx = eval(x.data(), g.data());
    pf[0] = fx;
    for (i = 0; i < n; ++i) d[i] = -g[i];
    gnorm_inf = inf_norm(g);
    xnorm_inf = inf_norm(x);
    k = 0;
    if (gnorm_inf / std::max(1.0, xnorm_inf) <= param.g_epsilon) {
        ret = LBFGS_CONVERGENCE;
    } else {
        step = 1.0 / std::sqrt(vdot(d, d));
        k = 1; end = 0; bound = 0;
        while (true) {
            xp = x; gp = g;
            step_min = param.min_step;
            step_max = param.max_step;
            step = step < step_max ? step : 0.5 * step_max;
            ls = line_search_lewisoverton(x, fx, g, step, d, xp, gp, step_min, step_max, eval, param);
            if (ls < 0) { x = xp; g = gp; ret = ls; break; }
            if (progress && progress(x.data(), g.data(), fx, step, k, ls)) { ret = LBFGS_CANCELED; break; }
            gnorm_inf = inf_norm(g);
            xnorm_inf = inf_norm(x);
            if (gnorm_inf / std::max(1.0, xnorm_inf) < param.g_epsilon) { ret = LBFGS_CONVERGENCE; break; }
            if (0 < param.past) {
                if (param.past <= k) {
                    rate = std::fabs(pf[k % param.past] - fx) / std::max(1.0, std::fabs(fx));
                    if (rate < param.delta) { ret = LBFGS_STOP; break; }
                }
                pf[k % param.past] = fx;
            }
            if (param.max_iterations != 0 && param.max_iterations <= k) { ret = LBFGSERR_MAXIMUMITERATION; break; }
            ++k;
            for (i = 0; i < n; ++i) { lm_s[end][i] = x[i] - xp[i]; lm_y[end][i] = g[i] - gp[i]; }
            ys = vdot(lm_y[end], lm_s[end]);
            yy = vdot(lm_y[end], lm_y[end]);
            lm_ys[end] = ys;
            for (i = 0; i < n; ++i) d[i] = -g[i];
            cau = vdot(lm_s[end], lm_s[end]) * std::sqrt(vdot(gp, gp)) * param.cautious_factor;
            if (ys > cau) {
                ++bound;
                bound = m < bound ? m : bound;
                end = (end + 1) % m;
                j = end;
                for (i = 0; i < bound; ++i) {
                    j = (j + m - 1) % m;
                    lm_alpha[j] = vdot(lm_s[j], d) / lm_ys[j];
                    for (int q = 0; q < n; ++q) d[q] += (-lm_alpha[j]) * lm_y[j][q];
                }
                for (int q = 0; q < n; ++q) d[q] *= ys / yy;
                for (i = 0; i < bound; ++i) {
                    beta = vdot(lm_y[j], d) / lm_ys[j];
                    for (int q = 0; q < n; ++q) d[q] += (lm_alpha[j] - beta) * lm_s[j][q];
                    j = (j + 1) % m;
                }
            }
            step = 1.0;
        }
    }
    f = fx;
    if (iters_out) *iters_out = k;
    return ret;
}

}  // namespace oracle
