// host/minco.hpp — host side of the drop-in: minimum-jerk (s = 3) non-uniform MINCO spline, its energy and the
// adjoint that maps d/d(coefficients, durations) back to d/d(waypoints, durations).
//
// Same mathematics and public surface as the reference's MINCO_S3NU
// (src/utils/include/utils/minco.hpp:397-655: setConditions / setParameters / getEnergy /
// getEnergyPartialGradByCoeffs / getEnergyPartialGradByTimes / propogateGrad), written without Eigen.
// The 6N x 6N system is banded (half-bandwidth 6); it is stored row-wise in a dense band and eliminated without
// pivoting, which the structure of the system permits (the reference does the same, minco.hpp:99-131).
// O(N) work, N <= 64: this stays on the host.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace svsdf {
namespace host {

class BandMatrix {
   public:
    static constexpr int BW = 6;
    static constexpr int W = 2 * BW + 1;
    void resize(int n) {
        n_ = n;
        a_.assign((size_t)n * W, 0.0);
    }
    void zero() { std::fill(a_.begin(), a_.end(), 0.0); }
    inline double &at(int r, int c) { return a_[(size_t)r * W + (c - r + BW)]; }
    inline double at(int r, int c) const { return a_[(size_t)r * W + (c - r + BW)]; }
    int n() const { return n_; }

    // In-place LU (Doolittle, unit lower), no pivoting.
    void factorize() {
        for (int k = 0; k < n_ - 1; ++k) {
            const double piv = at(k, k);
            const int rmax = std::min(k + BW, n_ - 1);
            const int cmax = std::min(k + BW, n_ - 1);
            for (int r = k + 1; r <= rmax; ++r) {
                double &l = at(r, k);
                if (l == 0.0) continue;
                l /= piv;
                for (int c = k + 1; c <= cmax; ++c) {
                    const double u = at(k, c);
                    if (u != 0.0) at(r, c) -= l * u;
                }
            }
        }
    }
    // Solve A X = B for m right-hand sides; B is n x m column-major (B[c*n + r]), overwritten by X.
    void solve(double *B, int m) const {
        for (int c = 0; c < m; ++c) {
            double *b = B + (size_t)c * n_;
            for (int j = 0; j < n_; ++j) {
                const int rmax = std::min(j + BW, n_ - 1);
                const double bj = b[j];
                for (int r = j + 1; r <= rmax; ++r) {
                    const double l = at(r, j);
                    if (l != 0.0) b[r] -= l * bj;
                }
            }
            for (int j = n_ - 1; j >= 0; --j) {
                b[j] /= at(j, j);
                const int rmin = std::max(0, j - BW);
                const double bj = b[j];
                for (int r = rmin; r < j; ++r) {
                    const double u = at(r, j);
                    if (u != 0.0) b[r] -= u * bj;
                }
            }
        }
    }
    // Solve A^T X = B.
    void solve_transposed(double *B, int m) const {
        for (int c = 0; c < m; ++c) {
            double *b = B + (size_t)c * n_;
            for (int j = 0; j < n_; ++j) {  // U^T is lower triangular
                b[j] /= at(j, j);
                const int cmax = std::min(j + BW, n_ - 1);
                const double bj = b[j];
                for (int i = j + 1; i <= cmax; ++i) {
                    const double u = at(j, i);
                    if (u != 0.0) b[i] -= u * bj;
                }
            }
            for (int j = n_ - 1; j >= 0; --j) {  // L^T is unit upper triangular
                const int cmin = std::max(0, j - BW);
                const double bj = b[j];
                for (int i = cmin; i < j; ++i) {
                    const double l = at(j, i);
                    if (l != 0.0) b[i] -= l * bj;
                }
            }
        }
    }

   private:
    int n_ = 0;
    std::vector<double> a_;
};

class MincoS3NU {
   public:
    // headState / tailState: 3x3 column-major (column k = k-th derivative), as the reference's Matrix3d.
    void setConditions(const double *headState, const double *tailState, int pieceNum) {
        N_ = pieceNum;
        std::memcpy(head_, headState, sizeof(head_));
        std::memcpy(tail_, tailState, sizeof(tail_));
        A_.resize(6 * N_);
        b_.assign((size_t)18 * N_, 0.0);
        t_.assign((size_t)5 * N_, 0.0);
    }
    int pieces() const { return N_; }
    // inPs: 3 x (N-1) column-major; ts: N durations.
    void setParameters(const double *inPs, const double *ts) {
        const int N = N_, n = 6 * N;
        for (int i = 0; i < N; ++i) {
            double *p = &t_[5 * i];
            p[0] = ts[i];
            p[1] = p[0] * p[0];
            p[2] = p[1] * p[0];
            p[3] = p[1] * p[1];
            p[4] = p[3] * p[0];
        }
        A_.zero();
        std::fill(b_.begin(), b_.end(), 0.0);
        auto B = [&](int r, int d) -> double & { return b_[(size_t)d * n + r]; };
        // head: position, velocity, acceleration
        A_.at(0, 0) = 1.0;
        A_.at(1, 1) = 1.0;
        A_.at(2, 2) = 2.0;
        for (int d = 0; d < 3; ++d)
            for (int k = 0; k < 3; ++k) B(k, d) = head_[k * 3 + d];
        for (int i = 0; i < N - 1; ++i) {
            const double *p = &t_[5 * i];
            const double T1 = p[0], T2 = p[1], T3 = p[2], T4 = p[3], T5 = p[4];
            const int r = 6 * i;
            // jerk, snap continuity
            A_.at(r + 3, r + 3) = 6.0;  A_.at(r + 3, r + 4) = 24.0 * T1;  A_.at(r + 3, r + 5) = 60.0 * T2;  A_.at(r + 3, r + 9) = -6.0;
            A_.at(r + 4, r + 4) = 24.0; A_.at(r + 4, r + 5) = 120.0 * T1; A_.at(r + 4, r + 10) = -24.0;
            // waypoint, position continuity
            const double pw[6] = {1.0, T1, T2, T3, T4, T5};
            for (int k = 0; k < 6; ++k) { A_.at(r + 5, r + k) = pw[k]; A_.at(r + 6, r + k) = pw[k]; }
            A_.at(r + 6, r + 6) = -1.0;
            // velocity, acceleration continuity
            A_.at(r + 7, r + 1) = 1.0; A_.at(r + 7, r + 2) = 2 * T1; A_.at(r + 7, r + 3) = 3 * T2; A_.at(r + 7, r + 4) = 4 * T3; A_.at(r + 7, r + 5) = 5 * T4; A_.at(r + 7, r + 7) = -1.0;
            A_.at(r + 8, r + 2) = 2.0; A_.at(r + 8, r + 3) = 6 * T1; A_.at(r + 8, r + 4) = 12 * T2; A_.at(r + 8, r + 5) = 20 * T3; A_.at(r + 8, r + 8) = -2.0;
            for (int d = 0; d < 3; ++d) B(r + 5, d) = inPs[i * 3 + d];
        }
        {
            const double *p = &t_[5 * (N - 1)];
            const double T1 = p[0], T2 = p[1], T3 = p[2], T4 = p[3], T5 = p[4];
            const int r = n - 6;
            const double pw[6] = {1.0, T1, T2, T3, T4, T5};
            for (int k = 0; k < 6; ++k) A_.at(n - 3, r + k) = pw[k];
            A_.at(n - 2, r + 1) = 1.0; A_.at(n - 2, r + 2) = 2 * T1; A_.at(n - 2, r + 3) = 3 * T2; A_.at(n - 2, r + 4) = 4 * T3; A_.at(n - 2, r + 5) = 5 * T4;
            A_.at(n - 1, r + 2) = 2.0; A_.at(n - 1, r + 3) = 6 * T1; A_.at(n - 1, r + 4) = 12 * T2; A_.at(n - 1, r + 5) = 20 * T3;
            for (int d = 0; d < 3; ++d)
                for (int k = 0; k < 3; ++k) B(n - 3 + k, d) = tail_[k * 3 + d];
        }
        A_.factorize();
        A_.solve(b_.data(), 3);
    }
    // MINCO coefficient matrix b: 6N x 3, column-major (offset d*6N + 6i + k)
    const double *getCoeffs() const { return b_.data(); }
    const double *durations5() const { return t_.data(); }

    double getEnergy() const {
        double e = 0.0;
        for (int i = 0; i < N_; ++i) {
            const double *p = &t_[5 * i];
            e += 36.0 * dot(i, 3, 3) * p[0] + 144.0 * dot(i, 4, 3) * p[1] + 192.0 * dot(i, 4, 4) * p[2] +
                 240.0 * dot(i, 5, 3) * p[2] + 720.0 * dot(i, 5, 4) * p[3] + 720.0 * dot(i, 5, 5) * p[4];
        }
        return e;
    }
    void getEnergyPartialGradByCoeffs(double *gdC) const {
        const int n = 6 * N_;
        for (int i = 0; i < N_; ++i) {
            const double *p = &t_[5 * i];
            for (int d = 0; d < 3; ++d) {
                const double *c = &b_[(size_t)d * n + 6 * i];
                double *g = gdC + (size_t)d * n + 6 * i;
                g[0] = g[1] = g[2] = 0.0;
                g[3] = 72.0 * c[3] * p[0] + 144.0 * c[4] * p[1] + 240.0 * c[5] * p[2];
                g[4] = 144.0 * c[3] * p[1] + 384.0 * c[4] * p[2] + 720.0 * c[5] * p[3];
                g[5] = 240.0 * c[3] * p[2] + 720.0 * c[4] * p[3] + 1440.0 * c[5] * p[4];
            }
        }
    }
    void getEnergyPartialGradByTimes(double *gdT) const {
        for (int i = 0; i < N_; ++i) {
            const double *p = &t_[5 * i];
            gdT[i] = 36.0 * dot(i, 3, 3) + 288.0 * dot(i, 4, 3) * p[0] + 576.0 * dot(i, 4, 4) * p[1] +
                     720.0 * dot(i, 5, 3) * p[1] + 2880.0 * dot(i, 5, 4) * p[2] + 3600.0 * dot(i, 5, 5) * p[3];
        }
    }
    // propogateGrad: gradByPoints 3 x (N-1) column-major, gradByTimes N.
    void propogateGrad(const double *partialGradByCoeffs, const double *partialGradByTimes, double *gradByPoints,
                       double *gradByTimes) const {
        const int N = N_, n = 6 * N;
        adj_.assign(partialGradByCoeffs, partialGradByCoeffs + (size_t)3 * n);
        A_.solve_transposed(adj_.data(), 3);
        for (int i = 0; i < N - 1; ++i)
            for (int d = 0; d < 3; ++d) gradByPoints[i * 3 + d] = adj_[(size_t)d * n + 6 * i + 5];
        for (int i = 0; i < N; ++i) {
            const double *p = &t_[5 * i];
            const double T1 = p[0], T2 = p[1], T3 = p[2], T4 = p[3];
            double s = 0.0;
            for (int d = 0; d < 3; ++d) {
                const double *c = &b_[(size_t)d * n + 6 * i];
                const double nvel = -(c[1] + 2.0 * T1 * c[2] + 3.0 * T2 * c[3] + 4.0 * T3 * c[4] + 5.0 * T4 * c[5]);
                const double nacc = -(2.0 * c[2] + 6.0 * T1 * c[3] + 12.0 * T2 * c[4] + 20.0 * T3 * c[5]);
                const double njer = -(6.0 * c[3] + 24.0 * T1 * c[4] + 60.0 * T2 * c[5]);
                if (i < N - 1) {
                    const double nsna = -(24.0 * c[4] + 120.0 * T1 * c[5]);
                    const double ncra = -120.0 * c[5];
                    const double *a = &adj_[(size_t)d * n + 6 * i + 3];
                    // rows 6i+3..6i+8: jerk, snap, waypoint, position, velocity, acceleration constraints
                    s += nsna * a[0] + ncra * a[1] + nvel * a[2] + nvel * a[3] + nacc * a[4] + njer * a[5];
                } else {
                    const double *a = &adj_[(size_t)d * n + n - 3];
                    s += nvel * a[0] + nacc * a[1] + njer * a[2];
                }
            }
            gradByTimes[i] = s + partialGradByTimes[i];
        }
    }

   private:
    double dot(int i, int r1, int r2) const {
        const int n = 6 * N_;
        double s = 0.0;
        for (int d = 0; d < 3; ++d) s += b_[(size_t)d * n + 6 * i + r1] * b_[(size_t)d * n + 6 * i + r2];
        return s;
    }
    int N_ = 0;
    double head_[9], tail_[9];
    BandMatrix A_;
    std::vector<double> b_, t_;
    mutable std::vector<double> adj_;
};

// tau <-> T diffeomorphism (back_end_optimizer.hpp:199-289)
inline double forwardT(double tau) {
    return tau > 0.0 ? ((0.5 * tau + 1.0) * tau + 1.0) : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0);
}
inline double backwardT(double T) { return T > 1.0 ? (std::sqrt(2.0 * T - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / T - 1.0)); }
inline double backwardGradT(double tau, double gradT) {
    if (tau > 0) return gradT * (tau + 1.0);
    const double den = (0.5 * tau - 1.0) * tau + 1.0;
    return gradT * (1.0 - tau) / (den * den);
}

}  // namespace host
}  // namespace svsdf
