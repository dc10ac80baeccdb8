"""GPU parity tests of the triangle-mesh shape functor (BasicShape::getonlySDF_igl,
Shape.hpp:332-340) through the C ABI: svsdf_config.mesh_* selects it, everything downstream (outer solve, GSIP branch,
cost + gradient, callback, optimiser) is the same code as for the analytic shapes.

Parity statement: the winding number is the reference's SINGLE-precision 4-way hierarchy with order-2 expansions; the host
builds it (csrc/host/fwn_bvh.hpp, bitwise the reference's tree and coefficients — tests/test_oracle_mesh.py) and the device
walks it with an explicit stack in the recursion's summation order, every float operation an un-fused *_rn intrinsic: the
value equals the reference's own compiled code BIT FOR BIT (tests/golden/fwn_ref.npz).  The closest-triangle distance is the
plain double minimum (box-pruned over the same tree).  Hence the functor value, its FD gradient and every per-point result
of the outer solve are BIT-IDENTICAL to the oracle's; sums differ by summation order only."""
import os

import numpy as np
import pytest

from implicit_svsdf_planner_b200 import api, scenes

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def nrel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def synthetic_mesh():
    return scenes.extrude_outline(scenes.star_outline(n_per_edge=2), half_height=0.49)


def reference_star_mesh():
    g = np.load(os.path.join(HERE, "golden", "fwn_ref.npz"))
    return g["star_V"], g["star_F"]


@pytest.mark.parametrize("which", ["synthetic", "reference_star_obj"])
def test_mesh_functor_is_bitwise_the_oracles(oracle_mod, which):
    m = synthetic_mesh() if which == "synthetic" else reference_star_mesh()
    rng = np.random.default_rng(31)
    rel = np.c_[rng.uniform(-4.5, 4.5, size=(20_000, 2)), np.zeros(20_000)]
    for pp in ((0.0, 0.0, 0.0), (0.6, -0.3, 25.0)):
        ctx = api.Context("ignored", poly_params=pp, strict_fp=True, mesh=m)
        s_g = ctx.shape_sdf(rel)
        s_c = oracle_mod.mesh_eval(m, rel, "sdf", poly_params=pp)
        bad = np.flatnonzero(s_g != s_c)
        assert bad.size == 0, (pp, bad.size, rel[bad[:3]], s_g[bad[:3]], s_c[bad[:3]])
        g_g = ctx.shape_grad1(rel)
        g_c = oracle_mod.mesh_eval(m, rel, "grad1", poly_params=pp)
        assert np.array_equal(g_g, g_c)
        ctx.close()
    # FMA-contracted build: same function to rounding
    ctx = api.Context("ignored", strict_fp=False, mesh=m)
    assert np.abs(ctx.shape_sdf(rel) - oracle_mod.mesh_eval(m, rel, "sdf")).max() < 1e-12
    ctx.close()


@pytest.mark.parametrize("name", ["star", "sdHorseshoe", "sdArc"])
def test_device_winding_number_is_bitwise_the_reference_golden(oracle_mod, name):
    """(1 - 2 w_ref) sqrt(d2) with w_ref from the REFERENCE's own compiled igl/HDK code on its own shapes/*.obj: the device
    functor returns exactly that, in both builds (the float traversal never contracts)."""
    g = np.load(os.path.join(HERE, "golden", "fwn_ref_sdarc.npz" if name == "sdArc" else "fwn_ref.npz"))  # sdArc: 2000 faces, 973 nodes
    m = (g[name + "_V"], g[name + "_F"])
    Q, w_ref = g[name + "_Q"][:2000], g[name + "_w_ref"][:2000]  # the z = 0 queries (the plane the planner evaluates in)
    assert np.all(Q[:, 2] == 0.0)
    d2 = oracle_mod.mesh_eval(m, Q, "sqr_distance")
    want = (1.0 - 2.0 * w_ref) * np.sqrt(d2)
    ctx = api.Context("ignored", strict_fp=True, mesh=m)
    got = ctx.shape_sdf(Q)
    assert np.array_equal(got, want), int((got != want).sum())
    ctx.close()
    ctx = api.Context("ignored", strict_fp=False, mesh=m)
    got = ctx.shape_sdf(Q)
    w_back = (1.0 - got / np.sqrt(d2)) / 2.0
    assert np.abs(w_back - w_ref).max() < 1e-12  # same float w; only the double distance may contract
    ctx.close()


def test_mesh_query_is_bit_identical_to_oracle(oracle_mod, scene_small_inside):
    m = synthetic_mesh()
    sc = scene_small_inside
    co = sc.coeffs_colmajor()
    opt = api.TrajOptimizer("ignored", strict_fp=True, mesh=m)
    sv = opt.sv_manager
    sv.updateTraj(sc.T, co)
    orc = oracle_mod.Oracle(mesh=m, threads=oracle_mod.num_procs())
    orc.set_traj(sc.T, co)
    p = np.c_[sc.points[:, :2], np.zeros(sc.P)]
    s_c, t_c, g_c = orc.query_outer(p)
    s_g, t_g, g_g = sv.getSDFofSweptVolume(p)
    assert np.array_equal(s_g, s_c) and np.array_equal(t_g, t_c) and np.array_equal(g_g, g_c)
    s_c, t_c, g_c, r_c = orc.query(p)
    s_g, t_g, g_g, r_g = sv.getTrueSDFofSweptVolume(p)
    assert np.array_equal(r_c, r_g) and (r_c > 0).sum() >= 5  # the GSIP branch is exercised
    outside = r_c == 0
    assert np.array_equal(s_g[outside], s_c[outside]) and np.array_equal(g_g[outside], g_c[outside])
    inside = ~outside
    assert np.array_equal(s_g[inside], s_c[inside]) and np.array_equal(g_g[inside], g_c[inside])


@pytest.mark.parametrize("which", ["synthetic", "reference_star_obj"])
def test_mesh_cost_grad_matches_oracle(oracle_mod, scene_small_inside, which):
    m = synthetic_mesh() if which == "synthetic" else reference_star_mesh()
    sc = scene_small_inside
    co = sc.coeffs_colmajor()
    opt = api.TrajOptimizer("ignored", weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, strict_fp=True, mesh=m)
    opt.parallel_points = sc.points
    orc = oracle_mod.Oracle(mesh=m, threads=oracle_mod.num_procs())
    orc.set_points(sc.points)
    c0, gT0, gC0, _, inside = orc.cost_grad(sc.T, co)
    c1, gT1, gC1 = opt.addSaftyPenaOnSweptVolumeParallelTrueSDF(sc.T, co)
    assert c0 > 0 and abs(c1 - c0) <= 1e-11 * abs(c0)
    assert nrel(gC1, gC0) <= 1e-8, nrel(gC1, gC0)
    assert np.linalg.norm(gT1 - gT0) <= 1e-8 * (np.linalg.norm(gT0) + 1e-3 * np.linalg.norm(gC0))
    # the FMA build stays inside the reference algorithm's own noise floor
    opt2 = api.TrajOptimizer("ignored", weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, strict_fp=False, mesh=m)
    opt2.parallel_points = sc.points
    c2, gT2, gC2 = opt2.addSaftyPenaOnSweptVolumeParallelTrueSDF(sc.T, co)
    assert abs(c2 - c0) <= 1e-8 * abs(c0) and nrel(gC2, gC0) <= 1e-3


def test_mesh_callback_and_optimiser(oracle_mod):
    m = synthetic_mesh()
    sc = scenes.make_scene("star", 6, 300, clearance=2.6)
    opt = api.TrajOptimizer("ignored", weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, strict_fp=True, mesh=m)
    opt.parallel_points = sc.points
    opt.setConditions(sc.init_s, sc.final_s, sc.N)
    orc = oracle_mod.Oracle(mesh=m, threads=oracle_mod.num_procs())
    orc.set_points(sc.points)
    orc.set_conditions(sc.init_s, sc.final_s, sc.N)
    f0, g0 = orc.evaluate(sc.x0)
    f1, g1 = opt.costFunction(sc.x0)
    assert abs(f1 - f0) <= 1e-11 * abs(f0) and nrel(g1, g0) <= 1e-8
    params = api.default_lbfgs_params(mem_size=16, past=3, delta=1e-5, g_epsilon=0.0, max_iterations=15, min_step=1e-32)
    rc, x, T, b, st = opt.optimize_traj(sc.init_s, sc.final_s, sc.x0, sc.N, params)
    assert st["final_cost"] < f1 and np.all(T > 0) and np.all(np.isfinite(b))


def test_mesh_config_errors():
    V, F = synthetic_mesh()
    bad = F.copy()
    bad[3, 1] = V.shape[0] + 7
    with pytest.raises(api.SvsdfError):
        api.Context("ignored", mesh=(V, bad))
    with pytest.raises(api.SvsdfError):
        api.Context("ignored", mesh=(V[:2], F))
