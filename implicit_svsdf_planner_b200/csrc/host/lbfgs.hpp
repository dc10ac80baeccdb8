// host/lbfgs.hpp — host outer optimiser of the drop-in: limited-memory BFGS with the Lewis–Overton weak-Wolfe
// line search (suitable for the piecewise-smooth SVSDF cost).  Algorithm and return codes follow the clean
// LBFGS-Lite the reference vendors (src/utils/include/utils/lbfgs_ref.hpp:276-395 line search, :434-716 driver);
// re-implemented without Eigen on plain arrays and re-entrant (no static state, unlike the reference's LMBM
// wrapper src/utils/include/utils/lmbm.cpp:4-6), so many problems can be optimised concurrently in batch mode.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace svsdf {
namespace host {

struct LbfgsParams {
    int mem_size = 8;
    int past = 3;
    double delta = 1.0e-6;
    double g_epsilon = 1.0e-5;
    int max_iterations = 0;
    int max_linesearch = 64;
    double min_step = 1.0e-20;
    double max_step = 1.0e+20;
    double f_dec_coeff = 1.0e-4;
    double s_curv_coeff = 0.9;
    double cautious_factor = 1.0e-6;
    double machine_prec = 1.0e-16;
    int nonsmooth_restarts = 0;   // see svsdf_lbfgs_params
    // 1: behave like the reference's patched utils/lbfgs.hpp (the solver of its mid end): the line search accepts on the Armijo
    // condition alone (the weak-Wolfe branch is `if (0)`, lbfgs.hpp:375) and a quasi-Newton direction with |d| >= 0.04 or d.g > 0 —
    // or a skipped cautious update — is replaced by -g scaled to the previous direction's length after one more evaluation at the
    // same point (lbfgs.hpp:759-779).
    int reference_patch = 0;
};

enum LbfgsCode {
    LBFGS_CONVERGENCE = 0,
    LBFGS_STOP = 1,
    LBFGS_CANCELED = 2,
    LBFGS_NOMOREPROGRESS = 3,  // line search failed along -g: no decrease to line-search precision (kink of a non-smooth cost)
    LBFGSERR_UNKNOWNERROR = -1024,
    LBFGSERR_INVALID_N,
    LBFGSERR_INVALID_MEMSIZE,
    LBFGSERR_INVALID_GEPSILON,
    LBFGSERR_INVALID_TESTPERIOD,
    LBFGSERR_INVALID_DELTA,
    LBFGSERR_INVALID_MINSTEP,
    LBFGSERR_INVALID_MAXSTEP,
    LBFGSERR_INVALID_FDECCOEFF,
    LBFGSERR_INVALID_SCURVCOEFF,
    LBFGSERR_INVALID_MACHINEPREC,
    LBFGSERR_INVALID_MAXLINESEARCH,
    LBFGSERR_INVALID_FUNCVAL,
    LBFGSERR_MINIMUMSTEP,
    LBFGSERR_MAXIMUMSTEP,
    LBFGSERR_MAXIMUMLINESEARCH,
    LBFGSERR_MAXIMUMITERATION,
    LBFGSERR_WIDTHTOOSMALL,
    LBFGSERR_INVALIDPARAMETERS,
    LBFGSERR_INCREASEGRADIENT,
};

typedef double (*lbfgs_eval_fn)(void *instance, const double *x, double *g, const int n);
typedef int (*lbfgs_progress_fn)(void *instance, const double *x, const int k);

struct LbfgsResult {
    double f = 0.0;
    int iterations = 0;
    int evaluations = 0;
    int status = 0;
};

class Lbfgs {
   public:
    explicit Lbfgs(const LbfgsParams &p) : P(p) {}

    LbfgsResult minimize(double *x, int n, lbfgs_eval_fn eval, void *inst, lbfgs_progress_fn progress, void *puser) {
        LbfgsResult R;
        if ((R.status = check(n)) != 0) return R;
        const int m = P.mem_size;
        xp.assign(n, 0.0); g.assign(n, 0.0); gp.assign(n, 0.0); d.assign(n, 0.0);
        pf.assign(std::max(1, P.past), 0.0);
        S.assign((size_t)m * n, 0.0); Y.assign((size_t)m * n, 0.0);
        ys_hist.assign(m, 0.0); alpha.assign(m, 0.0);

        double fx = eval(inst, x, g.data(), n);
        R.evaluations = 1;
        if (!std::isfinite(fx)) {  // a failed first evaluation (NaN cost, g left at zero) must not read as convergence
            R.status = LBFGSERR_INVALID_FUNCVAL;
            R.f = fx;
            return R;
        }
        pf[0] = fx;
        for (int i = 0; i < n; ++i) d[i] = -g[i];
        int k = 0;
        if (amax(g.data(), n) / std::max(1.0, amax(x, n)) <= P.g_epsilon) {
            R.status = LBFGS_CONVERGENCE;
        } else {
            double step = 1.0 / std::sqrt(dotp(d.data(), d.data(), n));
            int head = 0, stored = 0, restarts = 0;
            double olddnorm = 1.0;
            k = 1;
            for (;;) {
                std::copy(x, x + n, xp.begin());
                gp = g;
                const double step_min = P.min_step, step_max = P.max_step;
                step = step < step_max ? step : 0.5 * step_max;
                const double fx_before = fx;
                int ls = line_search(x, n, fx, step, step_min, step_max, eval, inst, R.evaluations);
                if (ls < 0) {
                    std::copy(xp.begin(), xp.end(), x);
                    g = gp;
                    fx = fx_before;
                    const bool recoverable = ls == LBFGSERR_MAXIMUMLINESEARCH || ls == LBFGSERR_MINIMUMSTEP || ls == LBFGSERR_WIDTHTOOSMALL ||
                                             ls == LBFGSERR_INCREASEGRADIENT || ls == LBFGSERR_MAXIMUMSTEP;
                    if (P.nonsmooth_restarts > 0 && recoverable) {
                        if (stored > 0 && restarts < P.nonsmooth_restarts) {
                            // the quasi-Newton model is wrong across a kink: forget it and search along -g from here
                            ++restarts;
                            stored = 0;
                            head = 0;
                            for (int i = 0; i < n; ++i) d[i] = -g[i];
                            step = 1.0 / std::sqrt(dotp(d.data(), d.data(), n));
                            continue;
                        }
                        R.status = LBFGS_NOMOREPROGRESS;
                        break;
                    }
                    R.status = ls;
                    break;
                }
                restarts = 0;
                if (progress && progress(puser, x, k)) { R.status = LBFGS_CANCELED; break; }
                if (amax(g.data(), n) / std::max(1.0, amax(x, n)) < P.g_epsilon) { R.status = LBFGS_CONVERGENCE; break; }
                if (P.past > 0) {
                    if (P.past <= k) {
                        const double rate = std::fabs(pf[k % P.past] - fx) / std::max(1.0, std::fabs(fx));
                        if (rate < P.delta) { R.status = LBFGS_STOP; break; }
                    }
                    pf[k % P.past] = fx;
                }
                if (P.max_iterations != 0 && P.max_iterations <= k) { R.status = LBFGSERR_MAXIMUMITERATION; break; }
                ++k;
                double *s = &S[(size_t)head * n], *y = &Y[(size_t)head * n];
                for (int i = 0; i < n; ++i) { s[i] = x[i] - xp[i]; y[i] = g[i] - gp[i]; }
                const double ys = dotp(y, s, n), yy = dotp(y, y, n);
                ys_hist[head] = ys;
                for (int i = 0; i < n; ++i) d[i] = -g[i];
                // cautious update (Li & Fukushima): skip the pair when curvature is too small
                const double cau = dotp(s, s, n) * std::sqrt(dotp(gp.data(), gp.data(), n)) * P.cautious_factor;
                if (ys > cau) {
                    stored = std::min(m, stored + 1);
                    head = (head + 1) % m;
                    int j = head;
                    for (int i = 0; i < stored; ++i) {  // two-loop recursion, newest to oldest
                        j = (j + m - 1) % m;
                        alpha[j] = dotp(&S[(size_t)j * n], d.data(), n) / ys_hist[j];
                        axpy(-alpha[j], &Y[(size_t)j * n], d.data(), n);
                    }
                    const double scale = ys / yy;
                    for (int i = 0; i < n; ++i) d[i] *= scale;
                    for (int i = 0; i < stored; ++i) {  // oldest to newest
                        const double beta = dotp(&Y[(size_t)j * n], d.data(), n) / ys_hist[j];
                        axpy(alpha[j] - beta, &S[(size_t)j * n], d.data(), n);
                        j = (j + 1) % m;
                    }
                    // (`!(|d| < 0.04)` rather than the reference's `|d| >= 0.04`: a non-finite direction — a stored pair with y.s == 0
                    // once the iterates stop moving — also takes the fallback instead of ending the run with a NaN cost)
                    if (P.reference_patch && (!(std::sqrt(dotp(d.data(), d.data(), n)) < 0.04) || dotp(d.data(), g.data(), n) > 0))
                        steepest_with_old_norm(x, n, eval, inst, R.evaluations, olddnorm);
                } else if (P.reference_patch) {
                    steepest_with_old_norm(x, n, eval, inst, R.evaluations, olddnorm);
                    head = (head + 1) % m;  // the slot just written is skipped without counting as stored (lbfgs.hpp:776)
                }
                if (P.reference_patch) olddnorm = std::sqrt(dotp(d.data(), d.data(), n));
                step = 1.0;
            }
        }
        R.f = fx;
        R.iterations = k;
        return R;
    }

   private:
    int check(int n) const {
        if (n <= 0) return LBFGSERR_INVALID_N;
        if (P.mem_size <= 0) return LBFGSERR_INVALID_MEMSIZE;
        if (P.g_epsilon < 0.0) return LBFGSERR_INVALID_GEPSILON;
        if (P.past < 0) return LBFGSERR_INVALID_TESTPERIOD;
        if (P.delta < 0.0) return LBFGSERR_INVALID_DELTA;
        if (P.min_step < 0.0) return LBFGSERR_INVALID_MINSTEP;
        if (P.max_step < P.min_step) return LBFGSERR_INVALID_MAXSTEP;
        if (!(P.f_dec_coeff > 0.0 && P.f_dec_coeff < 1.0)) return LBFGSERR_INVALID_FDECCOEFF;
        if (!(P.s_curv_coeff < 1.0 && P.s_curv_coeff > P.f_dec_coeff)) return LBFGSERR_INVALID_SCURVCOEFF;
        if (!(P.machine_prec > 0.0)) return LBFGSERR_INVALID_MACHINEPREC;
        if (P.max_linesearch <= 0) return LBFGSERR_INVALID_MAXLINESEARCH;
        return 0;
    }
    static double dotp(const double *a, const double *b, int n) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += a[i] * b[i];
        return s;
    }
    static void axpy(double a, const double *x, double *y, int n) {
        for (int i = 0; i < n; ++i) y[i] += a * x[i];
    }
    static double amax(const double *a, int n) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = std::max(r, std::fabs(a[i]));
        return r;
    }
    // the reference's fallback direction: re-evaluate at x (its own callback returns the same gradient again), d = -g / |g| * |d_prev|
    void steepest_with_old_norm(const double *x, int n, lbfgs_eval_fn eval, void *inst, int &evals, double olddnorm) {
        gp_scratch.assign(n, 0.0);
        (void)eval(inst, x, gp_scratch.data(), n);
        ++evals;
        g = gp_scratch;
        for (int i = 0; i < n; ++i) d[i] = -g[i];
        const double z = dotp(d.data(), d.data(), n);
        if (z > 0) {
            const double nn = std::sqrt(z);
            for (int i = 0; i < n; ++i) d[i] /= nn;
        }
        for (int i = 0; i < n; ++i) d[i] *= olddnorm;
    }
    // Lewis–Overton: Armijo + weak Wolfe by bisection / doubling.
    int line_search(double *x, int n, double &f, double &stp, double stpmin, double stpmax, lbfgs_eval_fn eval,
                    void *inst, int &evals) {
        if (!(stp > 0.0)) return LBFGSERR_INVALIDPARAMETERS;
        const double dginit = dotp(gp.data(), d.data(), n);
        if (0.0 < dginit) return LBFGSERR_INCREASEGRADIENT;
        const double finit = f, dgtest = P.f_dec_coeff * dginit, dstest = P.s_curv_coeff * dginit;
        double lo = 0.0, hi = stpmax;
        bool bracketed = false, touched = false;
        for (int count = 1;; ++count) {
            for (int i = 0; i < n; ++i) x[i] = xp[i] + stp * d[i];
            f = eval(inst, x, g.data(), n);
            ++evals;
            if (std::isinf(f) || std::isnan(f)) return LBFGSERR_INVALID_FUNCVAL;
            if (f > finit + stp * dgtest) {
                hi = stp;
                bracketed = true;
            } else if (!P.reference_patch && dotp(g.data(), d.data(), n) < dstest) {
                lo = stp;
            } else {
                return count;
            }
            if (P.max_linesearch <= count) return LBFGSERR_MAXIMUMLINESEARCH;
            if (bracketed && (hi - lo) < P.machine_prec * hi) return LBFGSERR_WIDTHTOOSMALL;
            stp = bracketed ? 0.5 * (lo + hi) : stp * 2.0;
            if (stp < stpmin) return LBFGSERR_MINIMUMSTEP;
            if (stp > stpmax) {
                if (touched) return LBFGSERR_MAXIMUMSTEP;
                touched = true;
                stp = stpmax;
            }
        }
    }

    LbfgsParams P;
    std::vector<double> xp, g, gp, d, pf, S, Y, ys_hist, alpha, gp_scratch;

};

}  // namespace host
}  // namespace svsdf
