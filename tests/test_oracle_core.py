"""CPU tests of the oracle's trajectory / swept-volume / cost / MINCO / L-BFGS restatement.

The control flow of the C++ oracle is cross-checked against small pure-Python restatements written from
SURVEY.md Appendix A (A.4 choiceTInit, A.5 gradientDescent, A.7 interior branch, A.8 penalty + chain rule) that use
the oracle only for the scalar SDF-at-time evaluation; analytic gradients are checked by finite differences."""
import math

import numpy as np
import pytest

from implicit_svsdf_planner_b200 import scenes

PI = 3.14159265358979323846


@pytest.fixture(scope="module")
def orc(oracle_mod, scene2k):
    o = oracle_mod.Oracle("star", threads=4)
    o.set_points(scene2k.points)
    o.set_traj(scene2k.T, scene2k.coeffs_colmajor())
    return o


def test_trajectory_eval_matches_numpy(orc, scene2k):
    sc = scene2k
    D = float(sc.T.sum())
    assert orc.duration() == D
    ts = np.linspace(0.0, D, 97)
    ref = scenes.eval_traj_xy(sc.coeffs, sc.T, ts)
    got = np.array([orc.traj_pos(t) for t in ts])
    assert np.abs(got - ref).max() < 1e-10
    # velocity by central differences of position
    for t in (0.3, 4.9, 11.2, 19.7):
        v = orc.traj_vel(t)
        fd = (orc.traj_pos(t + 1e-6) - orc.traj_pos(t - 1e-6)) / 2e-6
        assert np.abs(v - fd).max() < 1e-6
    # locatePieceIdx uses a strict '>' (trajectory.hpp:498-516): t == T0 is still evaluated on piece 0
    t = float(sc.T[0])
    c0 = sc.coeffs[:6]
    p_piece0 = sum(c0[k] * t**k for k in range(6))
    assert np.abs(orc.traj_pos(t) - p_piece0).max() < 1e-12
    # beyond the end: last piece extended (idx == N branch)
    c7 = sc.coeffs[-6:]
    tl = float(sc.T[-1]) + 0.25
    assert np.abs(orc.traj_pos(D + 0.25) - sum(c7[k] * tl**k for k in range(6))).max() < 1e-9


def py_choice_t_init(f, D, dt=0.15):
    """SURVEY.md A.4 / sw_manager.hpp:538-581."""
    min_dis, seed = 1e9, 0.0
    term = D
    t = 0.0
    for layer in range(1, 5):
        if layer == 1:
            t = 0.0
        else:
            t = max(0.0, seed - 10 * dt)
            term = min(D, seed + 10 * dt)
        while t <= term:
            d = f(t)
            if d < min_dis:
                seed, min_dis = t, d
            t += dt
        dt *= 0.1
    return seed


def py_gradient_descent(f, D, tmin, tmax, x0):
    """SURVEY.md A.5 / sw_manager.hpp:1249-1325 and :798-806."""
    def gdot(x):
        return (f(min(D, x + 0.000001)) - f(max(0.0, x - 0.000001))) * 500000

    x, prev, it, stop, fx = x0, 10000000.0, 0, False, None
    while it < 1000 and not stop and abs(x - prev) > 1e-16:
        if it == 0:
            fx = f(x)
        g = gdot(x)
        tau = 0.01
        prev = x
        for div in range(1, 30):
            it += 1
            g = gdot(x)
            change = -tau * (int(g > 0) - int(g < 0))
            xc = max(min(x + change, tmax), tmin)
            fc = f(xc)
            if fc - fx < 0:
                x, fx = xc, fc
                break
            tau = 0.5 * tau
            if div == 29:
                stop = True
    return fx, x


def test_choice_t_init_and_descent_match_python_restatement(orc, scene2k):
    D = orc.duration()
    rng = np.random.default_rng(5)
    idx = rng.choice(scene2k.P, size=12, replace=False)
    for i in idx:
        p = np.array([scene2k.points[i, 0], scene2k.points[i, 1], 0.0])
        f = lambda t: orc.sdf_at(p, t)
        seed_py = py_choice_t_init(f, D)
        seed_c = orc.choice_t_init(p, 0.15)
        assert seed_py == seed_c
        tmin, tmax = max(0.0, seed_c - 3.4), min(seed_c + 3.4, D)
        fx_py, x_py = py_gradient_descent(f, D, tmin, tmax, seed_c)
        fx_c, x_c = orc.gradient_descent(p, tmin, tmax, seed_c)
        assert (fx_py, x_py) == (fx_c, x_c)
        sdf, ts, g = orc.query_outer(p[None, :])
        assert sdf[0] == fx_c and ts[0] == x_c
        # t* is a local minimiser of t -> sdf(p, t) (up to the 3.7e-11 step floor and FD noise)
        assert f(x_c) <= f(min(D, x_c + 1e-5)) + 1e-12 and f(x_c) <= f(max(0.0, x_c - 1e-5)) + 1e-12


def test_eval_count_matches_survey_estimate(oracle_mod, scene2k):
    o = oracle_mod.Oracle("star", threads=1)
    o.set_traj(scene2k.T, scene2k.coeffs_colmajor())
    pts = np.c_[scene2k.points[::10, :2], np.zeros(200)]
    o.count_evals(True)
    o.query_outer(pts)
    per_point = o.eval_count() / 200
    # E = (floor(D/0.15)+1) + 63 + E_gd + 4, E_gd ~ 4e2 at D = 20 s
    assert 400 < per_point < 900, per_point


def py_true_sdf(orc, p, O):
    """SURVEY.md A.7 / sw_manager.hpp:916-1018 with the oracle's outer solve as the inner call (and the oracle's
    portable sin/cos/atan2, so the comparison with the C++ control flow is exact)."""
    _sc = lambda th: tuple(float(v[0]) for v in O.sincos([th]))  # (sin, cos)
    D = orc.duration()
    sdf, ts, g = orc.query_outer(np.array([[p[0], p[1], 0.0]]))
    if sdf[0] > 0:
        return sdf[0], ts[0], g[0], 0
    t_seed = ts[0]
    vel = orc.traj_vel(t_seed)
    if np.linalg.norm(vel) < 0.01:
        if t_seed < 0.1:
            t = t_seed
            while t <= D:
                vel = orc.traj_vel(t)
                if np.linalg.norm(vel) >= 0.01:
                    break
                t += 0.1
        elif t_seed > D - 0.1:
            t = t_seed
            while t >= 0:
                vel = orc.traj_vel(t)
                if np.linalg.norm(vel) >= 0.01:
                    break
                t -= 0.1
    r = 10.0
    theta0 = float(O.atan2([vel[0]], [-vel[1]])[0])
    if theta0 < 0:
        theta0 += 2 * PI
    theta_res = PI + 0.1
    it, rounds = 1, 0
    while True:
        max_g, star_theta, real_t = -100000, 0.0, 0.0
        th = theta0
        while th < theta0 + 2 * PI:
            sn, cs = _sc(th)
            y = np.array([[p[0] + 1.0 * r * cs, p[1] + 1.0 * r * sn, 0.0]])
            s, t, _ = orc.query_outer(y)
            if s[0] > max_g:
                max_g, real_t, star_theta = s[0], t[0], th
            th += theta_res
        r_star = r - max_g
        r = r_star
        rounds += 1
        if it > 8:
            break
        if abs(max_g) < 0.1:
            break
        theta_res = max(0.3, theta_res / 3)
        theta0 = star_theta
        it += 1
    sn, cs = _sc(star_theta)
    cor = np.array([p[0] + r_star * cs, p[1] + r_star * sn])
    gvec = np.array([cor[0] - p[0], cor[1] - p[1], 0.0])
    n = np.linalg.norm(gvec)
    if n > 0:
        gvec = gvec / n
    return -r_star, real_t, gvec, rounds


def test_interior_branch_matches_python_restatement(oracle_mod, scene_small_inside):
    sc = scene_small_inside
    o = oracle_mod.Oracle("star", threads=4)
    o.set_traj(sc.T, sc.coeffs_colmajor())
    pts = np.c_[sc.points[:, :2], np.zeros(sc.P)]
    sdf, ts, g, rounds = o.query(pts)
    inside = np.where(sdf <= 0)[0]
    assert inside.size >= 10
    assert np.all(rounds[sdf > 0] == 0) and np.all(rounds[inside] >= 1) and rounds.max() <= 9
    for i in inside[:6]:
        s_py, t_py, g_py, r_py = py_true_sdf(o, pts[i], oracle_mod)
        assert abs(s_py - sdf[i]) < 1e-12 and abs(t_py - ts[i]) < 1e-12 and r_py == rounds[i]
        assert np.abs(g_py - g[i]).max() < 1e-12
        assert abs(np.linalg.norm(g[i]) - 1.0) < 1e-12
    # GSIP result is (minus) a distance estimate to the swept-volume boundary: bounded by the shape's size
    assert sdf[inside].min() > -2.9


def py_smoothed_l1(x, mu=0.01):
    if x < 0:
        return None
    if x > mu:
        return x - 0.5 * mu, 1.0
    u = x / mu
    return (mu - 0.5 * x) * u * u * u, u * u * (-0.5 * u + 3.0 * (mu - 0.5 * x) / mu)


def test_penalty_and_chain_rule_match_python_restatement(oracle_mod, scene_small_inside):
    """SURVEY.md A.8 / back_end_optimizer.hpp:797-863, 1031-1066 recomputed in numpy from the per-point records."""
    sc = scene_small_inside
    o = oracle_mod.Oracle("star", threads=4)
    o.set_points(sc.points)
    co = sc.coeffs_colmajor()
    cost, gT, gC, pp, n_inside = o.cost_grad(sc.T, co, per_point=True)
    N = sc.N
    b = sc.coeffs  # 6N x 3
    starts = np.concatenate([[0.0], np.cumsum(sc.T)[:-1]])
    cost_py, gC_py, gT_py = 0.0, np.zeros((6 * N, 3)), np.zeros(N)
    for k in range(sc.P):
        sdf, tstar, gx, gy, _, piece, pena = pp[k]
        i = int(piece)
        s = tstar - starts[i]
        res = py_smoothed_l1(sc.safety_hor - sdf)
        if res is None or res[0] <= 0:
            assert pena == 0.0
            continue
        L, dL = res
        beta0 = np.array([s**q for q in range(6)])
        beta1 = np.array([q * s ** (q - 1) if q > 0 else 0.0 for q in range(6)])
        pos = beta0 @ b[6 * i : 6 * i + 6]
        vel = beta1 @ b[6 * i : 6 * i + 6]
        cy, sy = math.cos(pos[2]), math.sin(pos[2])
        R = np.array([[cy, -sy], [sy, cy]])
        g = np.array([gx, gy])  # body frame (already rotated for inside points)
        gxy = sc.weight_p * dL * (R @ g)
        d = sc.points[k, :2] - pos[:2]
        VRt = np.array([[-sy, cy], [-cy, -sy]])
        gyaw = -sc.weight_p * dL * (g @ (VRt @ d))
        G = np.array([gxy[0], gxy[1], gyaw])
        cost_py += sc.weight_p * L
        gC_py[6 * i : 6 * i + 6] += np.outer(beta0, G)
        gT_py[:i] += -(G @ vel)
        assert abs(pena - sc.weight_p * L) < 1e-9
    assert abs(cost_py - cost) / abs(cost) < 1e-12
    assert np.abs(gC_py.T.reshape(-1) - gC).max() / np.abs(gC).max() < 1e-9
    assert np.abs(gT_py - gT).max() / np.abs(gT).max() < 1e-9
    assert n_inside == int((pp[:, 0] <= 0).sum())


def test_cost_accumulates_into_inputs(oracle_mod, scene2k):
    sc = scene2k
    o = oracle_mod.Oracle("star", threads=4)
    o.set_points(sc.points[:300])
    co = sc.coeffs_colmajor()
    c0, gT0, gC0, _, _ = o.cost_grad(sc.T, co)
    rng = np.random.default_rng(0)
    aT, aC = rng.normal(size=sc.N), rng.normal(size=18 * sc.N)
    c1, gT1, gC1, _, _ = o.cost_grad(sc.T, co, cost0=5.0, gradT0=aT, gradC0=aC)
    assert abs(c1 - (c0 + 5.0)) < 1e-9 and np.allclose(gT1, gT0 + aT, atol=1e-9) and np.allclose(gC1, gC0 + aC, atol=1e-9)


def test_minco_matches_dense_numpy_solve(oracle_mod):
    for N, seed in ((2, 1), (8, 2), (16, 3)):
        init_s, final_s, q, T = scenes.make_trajectory("star", N, seed)
        T = T * np.random.default_rng(seed).uniform(0.6, 1.6, size=N)
        init_s[:, 1] = [0.3, -0.2, 0.1]
        final_s[:, 2] = [0.05, 0.02, -0.01]
        b_np = scenes.minco_dense(init_s, final_s, q, T)
        b, e, gdC, gdT = oracle_mod.minco_forward(init_s, final_s, q, T)
        assert np.abs(b - b_np).max() < 1e-9 * max(1.0, np.abs(b_np).max())
        # energy = integral of squared jerk (closed form minco.hpp:530-543) vs numerical quadrature
        D = T.sum()
        ts = np.linspace(0, D, 200001)
        starts = np.concatenate([[0.0], np.cumsum(T)[:-1]])
        idx = np.clip(np.searchsorted(np.cumsum(T), ts, side="left"), 0, N - 1)
        s = ts - starts[idx]
        c = b.reshape(N, 6, 3)[idx]
        jerk = 6 * c[:, 3] + 24 * c[:, 4] * s[:, None] + 60 * c[:, 5] * (s**2)[:, None]
        e_num = np.trapezoid((jerk**2).sum(axis=1), ts)
        assert abs(e - e_num) / e < 1e-6


def test_minco_adjoint_by_finite_differences(oracle_mod):
    N = 6
    init_s, final_s, q, T = scenes.make_trajectory("star", N, 9)
    rng = np.random.default_rng(9)
    T = T * rng.uniform(0.7, 1.4, size=N)
    W = rng.normal(size=(6 * N, 3))  # arbitrary linear functional of the coefficients
    wT = rng.normal(size=N)

    def J(qv, Tv):
        b, e, _, _ = oracle_mod.minco_forward(init_s, final_s, qv, Tv)
        return e + (W * b).sum() + wT @ Tv

    b, e, gdC, gdT = oracle_mod.minco_forward(init_s, final_s, q, T)
    gq, gT = oracle_mod.minco_propagate(init_s, final_s, q, T, gdC + W, gdT + wT)
    h = 1e-6
    for (d, i) in ((0, 0), (1, 2), (2, 4)):
        qp, qm = q.copy(), q.copy()
        qp[d, i] += h
        qm[d, i] -= h
        fd = (J(qp, T) - J(qm, T)) / (2 * h)
        assert abs(fd - gq[d, i]) <= 1e-5 * max(1.0, abs(fd)), (d, i, fd, gq[d, i])
    for i in (0, 3, N - 1):
        Tp, Tm = T.copy(), T.copy()
        Tp[i] += h
        Tm[i] -= h
        fd = (J(q, Tp) - J(q, Tm)) / (2 * h)
        assert abs(fd - gT[i]) <= 1e-5 * max(1.0, abs(fd)), (i, fd, gT[i])


def test_tau_maps_roundtrip(oracle_mod):
    T = np.array([0.2, 0.9, 1.0, 1.7, 2.5, 40.0])
    tau = scenes.backward_T(T)
    assert np.abs(scenes.forward_T(tau) - T).max() < 1e-12
    L = oracle_mod.lib()
    import ctypes as C

    out = np.empty_like(T)
    L.orc_backward_T(T.size, T.ctypes.data_as(oracle_mod.dp), out.ctypes.data_as(oracle_mod.dp))
    assert np.abs(out - tau).max() < 1e-15
    L.orc_forward_T(T.size, tau.ctypes.data_as(oracle_mod.dp), out.ctypes.data_as(oracle_mod.dp))
    assert np.abs(out - T).max() < 1e-12


def test_full_cost_gradient_by_directional_finite_differences(oracle_mod):
    """costFunctionLmbmParallel (back_end_optimizer.hpp:344-408): g must be the gradient of f.  The SVSDF term is
    piecewise smooth, so use a scene without interior points and a direction-averaged central difference."""
    sc = scenes.make_scene("star", 8, 300, clearance=3.0)
    o = oracle_mod.Oracle("star", threads=4)
    o.set_points(sc.points)
    o.set_conditions(sc.init_s, sc.final_s, sc.N)
    x0 = sc.x0
    f0, g0 = o.evaluate(x0)
    assert o.last_costs()[0] > 0  # the SVSDF penalty is active
    rng = np.random.default_rng(21)
    for _ in range(4):
        d = rng.normal(size=x0.size)
        d /= np.linalg.norm(d)
        h = 1e-5
        fp, _ = o.evaluate(x0 + h * d)
        fm, _ = o.evaluate(x0 - h * d)
        fd = (fp - fm) / (2 * h)
        assert abs(fd - g0 @ d) <= 2e-4 * max(1.0, abs(fd)), (fd, g0 @ d)
    # energy + time part alone (points far away): exact gradient
    far = sc.points.copy()
    far[:, 0] += 500.0
    o.set_points(far)
    f1, g1 = o.evaluate(x0)
    assert o.last_costs()[0] == 0.0
    d = np.zeros_like(x0)
    d[3] = 1.0
    fd = (o.evaluate(x0 + 1e-6 * d)[0] - o.evaluate(x0 - 1e-6 * d)[0]) / 2e-6
    assert abs(fd - g1[3]) < 1e-5 * max(1.0, abs(fd))


def test_lbfgs_on_rosenbrock_and_on_the_planner_cost(oracle_mod):
    import ctypes as C

    L = oracle_mod.lib()
    CB = C.CFUNCTYPE(C.c_double, C.c_void_p, oracle_mod.dp, oracle_mod.dp, C.c_int)

    def rosen(_i, xp, gp, n):
        x = np.ctypeslib.as_array(xp, shape=(n,))
        g = np.ctypeslib.as_array(gp, shape=(n,))
        g[0] = -400 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0])
        g[1] = 200 * (x[1] - x[0] ** 2)
        return 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2

    cb = CB(rosen)
    L.orc_lbfgs_cb.argtypes = [CB, C.c_void_p, oracle_mod.dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, oracle_mod.dp]
    x = np.array([-1.2, 1.0])
    stats = np.zeros(3)
    ret = L.orc_lbfgs_cb(cb, None, x.ctypes.data_as(oracle_mod.dp), 2, 8, 0, 1e-6, 1e-8, 0, stats.ctypes.data_as(oracle_mod.dp))
    assert ret == 0 and np.abs(x - 1.0).max() < 1e-6
    # planner cost: a few iterations reduce the cost and keep durations positive
    sc = scenes.make_scene("star", 8, 150, clearance=2.9)
    o = oracle_mod.Oracle("star", threads=4)
    o.set_points(sc.points)
    o.set_conditions(sc.init_s, sc.final_s, sc.N)
    f0, _ = o.evaluate(sc.x0)
    ret, x, st = o.lbfgs(sc.x0, mem_size=16, past=3, delta=1e-5, max_iterations=15)
    assert st["f"] < f0 and st["iters"] >= 1 and st["evals"] >= st["iters"]
    assert np.all(scenes.forward_T(x[: sc.N]) > 0)
