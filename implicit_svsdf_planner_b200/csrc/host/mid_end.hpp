// host/mid_end.hpp — the mid end: the warm-start optimisation that runs once per plan between the A* front end and the SVSDF
// back end.  Host code, no GPU: ~1 ms of work per plan.
//
// Restates, without Eigen / ROS,
//   OriTraj::getOriTraj                       src/planner_algorithm/src/mid_end.cpp:3-92
//   OriTraj::costFunction                     src/planner_algorithm/include/planner_algorithm/mid_end.hpp:277-325
//   OriTraj::addPosePenalty / grad_cost_dir   mid_end.hpp:183-273     cubic pull of every inner waypoint towards its A* cell
//   OriTraj::addTimeIntPenalty                mid_end.hpp:436-609     trapezoid integral of the velocity / body-rate / attitude penalties
//   OriTraj::smoothedL1, WC2, costaltitude, gradaltitude   mid_end.hpp:64-88, 374-434
//   flatness::FlatnessMap::forward / backward src/utils/include/utils/flatness.hpp:53-263  (multicopter differential flatness:
//                                             (v, a, j, psi, dpsi) -> thrust, attitude quaternion, body rate, and its adjoint)
// on top of this build's MINCO (host/minco.hpp) and L-BFGS (host/lbfgs.hpp).  The optimisation variable is the reference's:
// x = [tau (N), xi (3 (N - 1))], T = forwardT(tau), inner points = xi.
// Quirks of the reference that are kept because they define its numbers: costaltitude's `- 2 c1 (2 w x + y z)` term (the gradient
// routine differentiates `2 y z`), alpha = 0 in addPosePenalty (the pull acts at the START of piece i + 1, which is waypoint i),
// the attitude reference switching from the left to the right waypoint's rotation at the middle of a piece with the WC2 window.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "lbfgs.hpp"
#include "minco.hpp"

namespace svsdf {
namespace host {

struct MidEndConfig {
    // config/*.yaml keys read by OriTraj::setParam (mid_end.hpp:333-359); defaults = config/star.yaml
    double rho_mid_end = 2.0, vmax = 10.0, omgmax = 10.0, weight_v = 10.0, weight_omg = 10.0, weight_pr = 40.0, weight_ar = 0.0;
    double smoothingEps = 1.0e-2;
    int integralIntervs = 4;
    double vehicleMass = 0.61, gravAcc = 9.8, horizDrag = 0.10, vertDrag = 0.10, parasDrag = 0.01, speedEps = 0.0001;
    // L-BFGS (mid_end.cpp:46-52)
    int mem_size = 16, past = 64;
    double min_step = 1.0e-32, g_epsilon = 0.0, relCostTolMidEnd = 1.0e-10;
    int max_iterations = 10000, cancel_after = 100;  // earlyExit: `return k > 1e2` (mid_end.hpp:610-627)
    int solver = 0;  // 0: the reference's patched L-BFGS behaviour (same warm start as the reference); 1: this build's L-BFGS
};

// flatness.hpp:53-263.  forward() keeps its intermediates for backward(), exactly like the reference object.
class FlatMap {
   public:
    void reset(double m, double g, double horiz, double vert, double paras, double eps) { mass = m; grav = g; dh = horiz; dv = vert; cp = paras; veps = eps; }
    void forward(const double v[3], const double a[3], const double j[3], double psi, double dpsi, double &thr, double q[4], double w[3]) {
        vx = v[0]; vy = v[1]; vz = v[2]; ax = a[0]; ay = a[1]; az = a[2];
        speed = std::sqrt(vx * vx + vy * vy + vz * vz + veps);
        drag = 1.0 + cp * speed;
        const double wx0 = drag * vx, wy0 = drag * vy, wz0 = drag * vz;
        k_dh = dh / mass;
        ux = ax + k_dh * wx0; uy = ay + k_dh * wy0; uz = az + k_dh * wz0 + grav;
        uxx = ux * ux; uyy = uy * uy; uzz = uz * uz; uxy = ux * uy; uyz = uy * uz; uxz = ux * uz;
        un2 = uxx + uyy + uzz;
        un = std::sqrt(un2);
        zx = ux / un; zy = uy / un; zz = uz / un;
        un3 = un2 * un;
        p00 = (uyy + uzz) / un3; p01 = -uxy / un3; p02 = -uxz / un3; p11 = (uxx + uzz) / un3; p12 = -uyz / un3; p22 = (uxx + uyy) / un3;
        v_a = vx * ax + vy * ay + vz * az;
        ddrag = cp * v_a / speed;
        const double dwx = drag * ax + ddrag * vx, dwy = drag * ay + ddrag * vy, dwz = drag * az + ddrag * vz;
        ex = j[0] + k_dh * dwx; ey = j[1] + k_dh * dwy; ez = j[2] + k_dh * dwz;
        dzx = p00 * ex + p01 * ey + p02 * ez;
        dzy = p01 * ex + p11 * ey + p12 * ez;
        dzz = p02 * ex + p12 * ey + p22 * ez;
        fx = mass * ax + dv * wx0; fy = mass * ay + dv * wy0; fz = mass * (az + grav) + dv * wz0;
        thr = zx * fx + zy * fy + zz * fz;
        tden = std::sqrt(2.0 * (1.0 + zz));
        t0 = 0.5 * tden; t1 = -zy / tden; t2 = zx / tden;
        ch = std::cos(0.5 * psi); sh = std::sin(0.5 * psi);
        q[0] = t0 * ch; q[1] = t1 * ch + t2 * sh; q[2] = t2 * ch - t1 * sh; q[3] = t0 * sh;
        cps = std::cos(psi); sps = std::sin(psi);
        oden = zz + 1.0;
        oterm = dzz / oden;
        w[0] = dzx * sps - dzy * cps - (zx * sps - zy * cps) * oterm;
        w[1] = dzx * cps + dzy * sps - (zx * cps + zy * sps) * oterm;
        w[2] = (zy * dzx - zx * dzy) / oden + dpsi;
    }
    // adjoint of forward(): gradients w.r.t. (pos, vel, thr, quat, omg) -> total gradients w.r.t. (pos, vel, acc, jer, psi, dpsi)
    void backward(const double gp[3], const double gv[3], double gthr, const double gq[4], const double gw[3], double tp[3], double tv[3],
                  double ta[3], double tj[3], double &tpsi, double &tdpsi) const {
        const double t0b = sh * gq[3] + ch * gq[0];
        const double h3b = t0 * gq[3] + t2 * gq[1] - t1 * gq[2];
        const double t2b = ch * gq[2] + sh * gq[1];
        const double h0b = t2 * gq[2] + t1 * gq[1] + t0 * gq[0];
        const double t1b = ch * gq[1] - sh * gq[2];
        const double tden2 = tden * tden;
        const double tdenb = (zy * t1b - zx * t2b) / tden2 + 0.5 * t0b;
        const double otermb = -((zx * cps + zy * sps) * gw[1]) - (zx * sps - zy * cps) * gw[0];
        double tmp = gw[2] / oden;
        tdpsi = gw[2];
        double zyb = dzx * tmp;
        const double dzxb = zy * tmp + cps * gw[1] + sps * gw[0];
        double zxb = -(dzy * tmp);
        const double dzyb = sps * gw[1] - zx * tmp - cps * gw[0];
        const double odenb = -((zy * dzx - zx * dzy) * tmp / oden) - dzz * otermb / (oden * oden);
        tmp = -(oterm * gw[1]);
        double cpsb = dzx * gw[1] + zx * tmp;
        double spsb = dzy * gw[1] + zy * tmp;
        zxb += cps * tmp;
        zyb += sps * tmp;
        tmp = -(oterm * gw[0]);
        spsb += dzx * gw[0] + zx * tmp;
        cpsb += -dzy * gw[0] - zy * tmp;
        zxb += sps * tmp + t2b / tden + fx * gthr;
        zyb += -cps * tmp - t1b / tden + fy * gthr;
        const double dzzb = otermb / oden;
        const double zzb = odenb + tdenb / tden + fz * gthr;
        tpsi = cps * spsb + 0.5 * ch * h3b - sps * cpsb - 0.5 * sh * h0b;
        const double fxb = zx * gthr, fyb = zy * gthr, fzb = zz * gthr;
        const double p02b = ex * dzzb + ez * dzxb;
        const double exb = p02 * dzzb + p01 * dzyb + p00 * dzxb;
        const double p12b = ey * dzzb + ez * dzyb;
        const double eyb = p12 * dzzb + p11 * dzyb + p01 * dzxb;
        const double p22b = ez * dzzb;
        const double ezb = p22 * dzzb + p12 * dzyb + p02 * dzxb;
        const double p01b = ex * dzyb + ey * dzxb;
        const double p11b = ey * dzyb;
        const double p00b = ex * dzxb;
        tj[2] = ezb;
        const double dwzb = k_dh * ezb;
        tj[1] = eyb;
        const double dwyb = k_dh * eyb;
        tj[0] = exb;
        const double dwxb = k_dh * exb;
        tmp = cp * (vz * dwzb + vy * dwyb + vx * dwxb) / speed;
        ta[2] = mass * fzb + drag * dwzb + vz * tmp;
        ta[1] = mass * fyb + drag * dwyb + vy * tmp;
        ta[0] = mass * fxb + drag * dwxb + vx * tmp;
        tv[2] = ddrag * dwzb + az * tmp;
        tv[1] = ddrag * dwyb + ay * tmp;
        tv[0] = ddrag * dwxb + ax * tmp;
        double speedb = -(v_a * tmp / speed);
        tmp = p22b / un3;
        double uxxb = tmp, uyyb = tmp;
        double un3b = -((uxx + uyy) * tmp / un3);
        const double uyzb = -(p12b / un3);
        tmp = p11b / un3;
        un3b += uyz * p12b / (un3 * un3) - (uxx + uzz) * tmp / un3;
        uxxb += tmp;
        double uzzb = tmp;
        const double uxzb = -(p02b / un3);
        const double uxyb = -(p01b / un3);
        tmp = p00b / un3;
        un3b += uxz * p02b / (un3 * un3) + uxy * p01b / (un3 * un3) - (uyy + uzz) * tmp / un3;
        const double unb = un2 * un3b - (uz * zzb + uy * zyb + ux * zxb) / un2;
        const double un2b = un * un3b + unb / (2.0 * un);
        tmp += un2b;
        uyyb += tmp;
        uzzb += tmp;
        const double uzb = zzb / un + ux * uxzb + uy * uyzb + 2 * uz * uzzb;
        const double wzb = dv * fzb + k_dh * uzb;
        const double uyb = zyb / un + uz * uyzb + ux * uxyb + 2 * uy * uyyb;
        const double wyb = dv * fyb + k_dh * uyb;
        uxxb += un2b;
        const double uxb = zxb / un + uz * uxzb + uy * uxyb + 2 * ux * uxxb;
        const double wxb = dv * fxb + k_dh * uxb;
        const double dragb = az * dwzb + ay * dwyb + ax * dwxb + vz * wzb + vy * wyb + vx * wxb;
        ta[2] += uzb;
        ta[1] += uyb;
        ta[0] += uxb;
        speedb += cp * dragb;
        const double v2b = speedb / (2.0 * speed);
        tv[2] += drag * wzb + 2 * vz * v2b + gv[2];
        tv[1] += drag * wyb + 2 * vy * v2b + gv[1];
        tv[0] += drag * wxb + 2 * vx * v2b + gv[0];
        tp[2] = gp[2];
        tp[1] = gp[1];
        tp[0] = gp[0];
    }

   private:
    double mass = 1, grav = 9.8, dh = 0, dv = 0, cp = 0, veps = 1e-4;
    double vx, vy, vz, ax, ay, az, v_a, zx, zy, zz, dzx, dzy, dzz, speed, drag, k_dh, un2, un, ux, uy, uz, uxx, uyy, uzz, uxy, uyz, uxz;
    double p00, p01, p02, p11, p12, p22, un3, ddrag, ex, ey, ez, fx, fy, fz, tden, t0, t1, t2, ch, sh, cps, sps, oden, oterm;
};

class MidEnd {
   public:
    explicit MidEnd(const MidEndConfig &c) : cfg(c) { flat.reset(c.vehicleMass, c.gravAcc, c.horizDrag, c.vertDrag, c.parasDrag, c.speedEps); }

    // initS / finalS: 3x3 column-major; Q: 3 x (N - 1) column-major inner waypoints; rot: (N - 1) rotation matrices, 3x3 column-major each
    void setup(const double *initS, const double *finalS, int N_, const double *Q, const double *rot) {
        N = N_;
        minco.setConditions(initS, finalS, N);
        ref.assign(Q, Q + 3 * (size_t)(N - 1));
        att.assign(rot, rot + 9 * (size_t)(N - 1));
        T.assign(N, 0.0);
        P.assign(3 * (size_t)(N - 1), 0.0);
        gC.assign(18 * (size_t)N, 0.0);
        gT.assign(N, 0.0);
        gP.assign(3 * (size_t)(N - 1), 0.0);
        gTt.assign(N, 0.0);
    }
    int dim() const { return N + 3 * (N - 1); }

    static bool smoothedL1(double x, double mu, double &f, double &df) {  // mid_end.hpp:64-88
        if (x < 0.0) return false;
        if (x > mu) { f = x - 0.5 * mu; df = 1.0; return true; }
        const double xdmu = x / mu, sq = xdmu * xdmu, m = mu - 0.5 * x;
        f = m * sq * xdmu;
        df = sq * ((-0.5) * xdmu + 3.0 * m / mu);
        return true;
    }
    static double WC2(double x, double &dx) {  // mid_end.hpp:418-434
        if (x < -1) { dx = 0; return 0; }
        if (x < -0.5) { dx = 4 * (x + 1); return 2 * (x + 1) * (x + 1); }
        if (x < 0.5) { dx = -4 * x; return 1 - 2 * x * x; }
        if (x < 1) { dx = 4 * (x - 1); return 2 * (x - 1) * (x - 1); }
        dx = 0;
        return 0;
    }
    // R: 3x3 column-major (R(r, c) = R[3 c + r])
    static double cost_attitude(const double q[4], const double *R) {  // costaltitude, mid_end.hpp:374-391
        const double w = q[0], x = q[1], y = q[2], z = q[3];
        const double a0 = R[0], a1 = R[3], a2 = R[6], b0 = R[1], b1 = R[4], b2 = R[7], c0 = R[2], c1 = R[5], c2 = R[8];
        return 2 * a0 * (2 * y * y + 2 * z * z - 1) + 2 * b1 * (2 * x * x + 2 * z * z - 1) + 2 * c2 * (2 * x * x + 2 * y * y - 1) + 2 * a1 * (2 * w * z - 2 * x * y) -
               2 * a2 * (2 * w * y + 2 * x * z) - 2 * b0 * (2 * w * z + 2 * x * y) + 2 * b2 * (2 * w * x - 2 * y * z) + 2 * c0 * (2 * w * y - 2 * x * z) -
               2 * c1 * (2 * w * x + y * z) + 6;
    }
    static void grad_attitude(const double q[4], const double *R, double d[4]) {  // gradaltitude, mid_end.hpp:393-416
        const double w = q[0], x = q[1], y = q[2], z = q[3];
        const double a0 = R[0], a1 = R[3], a2 = R[6], b0 = R[1], b1 = R[4], b2 = R[7], c0 = R[2], c1 = R[5], c2 = R[8];
        d[0] = 4 * (b2 * x - a2 * y + a1 * z - c1 * x - b0 * z + c0 * y);
        d[1] = 4 * (b2 * w - a1 * y + 2 * b1 * x - c1 * w - b0 * y - a2 * z + 2 * c2 * x - c0 * z);
        d[2] = 4 * (2 * a0 * y - a1 * x - a2 * w - b0 * x + c0 * w - b2 * z + 2 * c2 * y - c1 * z);
        d[3] = 4 * (a1 * w - b0 * w - a2 * x + 2 * a0 * z - c0 * x - b2 * y + 2 * b1 * z - c1 * y);
    }

    // costFunction (mid_end.hpp:277-325): x = [tau, xi] -> cost, g
    double cost(const double *x, double *g) {
        const int n = 6 * N;
        for (int i = 0; i < N; ++i) T[i] = forwardT(x[i]);
        for (size_t k = 0; k < P.size(); ++k) P[k] = x[N + k];
        minco.setParameters(P.data(), T.data());
        double c = minco.getEnergy();
        minco.getEnergyPartialGradByCoeffs(gC.data());
        minco.getEnergyPartialGradByTimes(gT.data());
        const double *b = minco.getCoeffs();
        auto row = [&](int i, const double beta[6], double out[3]) {  // c^T beta of piece i
            for (int d = 0; d < 3; ++d) {
                const double *cc = b + (size_t)d * n + 6 * i;
                double s = cc[0] * beta[0];
                for (int k = 1; k < 6; ++k) s += cc[k] * beta[k];
                out[d] = s;
            }
        };
        // ---- addPosePenalty (:213-273): alpha = 0, i.e. the start of piece i + 1 ----
        for (int i = 0; i < N - 1; ++i) {
            const int seg = i + 1;
            const double b0[6] = {1.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            double pos[3];
            row(seg, b0, pos);
            const double dx = pos[0] - ref[3 * i], dy = pos[1] - ref[3 * i + 1], dz = pos[2] - ref[3 * i + 2];
            const double nrm = std::sqrt(dx * dx + dy * dy + dz * dz);
            const double cp3 = std::pow(nrm, 3);                 // grad_cost_dir (:183-210)
            if (cp3 + 0.0 > 0) {
                const double k3 = 3 * std::pow(nrm, 2);
                const double gpv[3] = {k3 * (dx / nrm), k3 * (dy / nrm), k3 * (dz / nrm)};  // 3 |d|^2 * d.normalized()
                for (int d = 0; d < 3; ++d) gC[(size_t)d * n + 6 * seg] += cfg.weight_pr * (1.0 * gpv[d]);  // beta0 = (1, 0, 0, 0, 0, 0)
                gT[seg] += cfg.weight_pr * (cp3 * 0.0);          // gradViolaPt = alpha * grad.dot(vel), alpha = 0 (keeps NaN / inf semantics out: finite operands)
                c += cfg.weight_pr * cp3 + cfg.weight_ar * 0.0;
            }
        }
        // ---- addTimeIntPenalty (:436-609) ----
        const double vmax2 = cfg.vmax * cfg.vmax, omax2 = cfg.omgmax * cfg.omgmax;
        const int res = cfg.integralIntervs;
        const double frac = 1.0 / res;
        const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int i = 0; i < N; ++i) {
            const double step = T[i] * frac, midT = T[i] * 0.5;
            const double *RL = (i > 0) ? &att[9 * (size_t)(i - 1)] : I3;
            const double *RR = (i < N - 1) ? &att[9 * (size_t)i] : I3;
            for (int j = 0; j <= res; ++j) {
                const double s1 = j * step, s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
                const double be0[6] = {1.0, s1, s2, s3, s4, s5};
                const double be1[6] = {0.0, 1.0, 2.0 * s1, 3.0 * s2, 4.0 * s3, 5.0 * s4};
                const double be2[6] = {0.0, 0.0, 2.0, 6.0 * s1, 12.0 * s2, 20.0 * s3};
                const double be3[6] = {0.0, 0.0, 0.0, 6.0, 24.0 * s1, 60.0 * s2};
                const double be4[6] = {0.0, 0.0, 0.0, 0.0, 24.0, 120.0 * s1};
                double vel[3], acc[3], jer[3], sna[3];
                row(i, be1, vel); row(i, be2, acc); row(i, be3, jer); row(i, be4, sna);
                double thr, quat[4], omg[3];
                flat.forward(vel, acc, jer, 0.0, 0.0, thr, quat, omg);
                const double *Rref;
                double dk, kRT;
                if (s1 <= midT) { Rref = RL; kRT = WC2(s1 / midT, dk); }
                else { Rref = RR; kRT = WC2((s1 - midT) / midT - 1.0, dk); }
                double pena = 0.0;
                double gVel[3] = {0, 0, 0}, gOmg[3] = {0, 0, 0}, gQuat[4] = {0, 0, 0, 0};
                const double gPos[3] = {0, 0, 0};
                const double catt = cost_attitude(quat, Rref);
                double datt[4];
                grad_attitude(quat, Rref, datt);
                const double violaVel = (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]) - vmax2;
                const double violaOmg = (omg[0] * omg[0] + omg[1] * omg[1] + omg[2] * omg[2]) - omax2;
                double f, df;
                if (smoothedL1(violaVel, cfg.smoothingEps, f, df)) {
                    for (int d = 0; d < 3; ++d) gVel[d] += cfg.weight_v * df * 2.0 * vel[d];
                    pena += cfg.weight_v * f;
                }
                if (smoothedL1(violaOmg, cfg.smoothingEps, f, df)) {
                    for (int d = 0; d < 3; ++d) gOmg[d] += cfg.weight_omg * df * 2.0 * omg[d];
                    pena += cfg.weight_omg * f;
                }
                if (smoothedL1(catt, cfg.smoothingEps, f, df)) {
                    for (int d = 0; d < 4; ++d) gQuat[d] += kRT * cfg.weight_ar * df * datt[d];
                    pena += kRT * cfg.weight_ar * f;
                }
                double tP[3], tV[3], tA[3], tJ[3], tpsi, tdpsi;
                flat.backward(gPos, gVel, 0.0, gQuat, gOmg, tP, tV, tA, tJ, tpsi, tdpsi);
                const double node = (j == 0 || j == res) ? 0.5 : 1.0;
                const double alpha = j * frac;
                for (int d = 0; d < 3; ++d)
                    for (int k = 0; k < 6; ++k)
                        gC[(size_t)d * n + 6 * i + k] += (be0[k] * tP[d] + be1[k] * tV[d] + be2[k] * tA[d] + be3[k] * tJ[d]) * node * step;
                const double dotp = tP[0] * vel[0] + tP[1] * vel[1] + tP[2] * vel[2], dotv = tV[0] * acc[0] + tV[1] * acc[1] + tV[2] * acc[2];
                const double dota = tA[0] * jer[0] + tA[1] * jer[1] + tA[2] * jer[2], dotj = tJ[0] * sna[0] + tJ[1] * sna[1] + tJ[2] * sna[2];
                gT[i] += (dotp + dotv + dota + dotj) * alpha * node * step + node * frac * pena;
                c += node * step * pena;
            }
        }
        minco.propogateGrad(gC.data(), gT.data(), gP.data(), gTt.data());
        double sumT = 0.0;
        for (int i = 0; i < N; ++i) sumT += T[i];
        c += cfg.rho_mid_end * sumT;
        for (int i = 0; i < N; ++i) g[i] = backwardGradT(x[i], gTt[i] + cfg.rho_mid_end);
        for (size_t k = 0; k < gP.size(); ++k) g[N + k] = gP[k];
        return c;
    }

    // getOriTraj (mid_end.cpp:3-92): x0 = [backwardT(T0), Q]; returns the L-BFGS status (>= 0: success; 2 = cancelled by the
    // `k > 100` rule like the reference's earlyExit); x_out = opt_x, T_out / coeffs_out the resulting spline.
    int optimize(const double *T0, double *x_out, double *T_out, double *coeffs_out, double *final_cost, int *iterations) {
        const int n = dim();
        std::vector<double> x(n);
        for (int i = 0; i < N; ++i) x[i] = backwardT(T0[i]);
        for (size_t k = 0; k < ref.size(); ++k) x[N + k] = ref[k];
        LbfgsParams lp;
        lp.mem_size = cfg.mem_size;
        lp.past = cfg.past;
        lp.min_step = cfg.min_step;
        lp.g_epsilon = cfg.g_epsilon;
        lp.delta = cfg.relCostTolMidEnd;
        lp.max_iterations = cfg.max_iterations;
        lp.reference_patch = cfg.solver == 0 ? 1 : 0;
        lp.nonsmooth_restarts = cfg.solver == 0 ? 0 : 8;  // own solver: a failed line search restarts from steepest descent, then status 3
        Lbfgs solver(lp);
        struct Hook { MidEnd *self; int cancel_after; int iters; } hook{this, cfg.cancel_after, 0};
        const LbfgsResult R = solver.minimize(
            x.data(), n, [](void *u, const double *xx, double *gg, const int) { return static_cast<Hook *>(u)->self->cost(xx, gg); }, &hook,
            [](void *u, const double *, const int k) { Hook *h = static_cast<Hook *>(u); h->iters = k; return (k > h->cancel_after) ? 1 : 0; }, &hook);
        const int ret = R.status;
        if (final_cost) *final_cost = R.f;
        if (iterations) *iterations = R.iterations;
        for (int i = 0; i < N; ++i) T[i] = forwardT(x[i]);
        for (size_t k = 0; k < P.size(); ++k) P[k] = x[N + k];
        minco.setParameters(P.data(), T.data());
        if (x_out) std::memcpy(x_out, x.data(), sizeof(double) * n);
        if (T_out) std::memcpy(T_out, T.data(), sizeof(double) * N);
        if (coeffs_out) std::memcpy(coeffs_out, minco.getCoeffs(), sizeof(double) * 18 * (size_t)N);
        return ret;
    }

   private:
    MidEndConfig cfg;
    FlatMap flat;
    MincoS3NU minco;
    int N = 0;
    std::vector<double> ref, att, T, P, gC, gT, gP, gTt;
};

}  // namespace host
}  // namespace svsdf
