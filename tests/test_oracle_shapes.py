"""CPU tests that pin the oracle's shape-SDF functors (oracle/shapes.hpp, restating
/root/reference/src/utils/include/utils/Shape.hpp) with (i) the reference's own shape meshes
(tests/golden/obj_outlines.json, generated from src/plan_manager/shapes/*.obj), (ii) closed-form known answers,
(iii) the Eikonal property of exact SDFs, (iv) the body-frame pre-transform."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ANALYTIC = ["star", "sdHorseshoe", "sdPie", "sdPie2", "sdArc", "sdTunnel", "sdCutDisk", "sdTrapezoid", "sdRhombus",
            "sdHeart", "sdRoundedX", "bigX", "sdRoundedCross", "sdOrientedVesica", "sdMoon", "sdUnevenCapsule"]


def _rel(xy):
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    return np.c_[xy, np.zeros(xy.shape[0])]


@pytest.fixture(scope="module")
def outlines():
    with open(os.path.join(HERE, "golden", "obj_outlines.json")) as fh:
        return json.load(fh)


# The .obj files are decimated Meshlab exports (star tip at 2.758 vs analytic 2.8), so the outline
# lies within a few cm of the analytic zero level set.  sdCutDisk.obj is visibly a different cut (its outline is 0.3-0.4
# inside the analytic shape and encloses 15 % less area) — a discrepancy inside the reference itself (the mesh is only
# displayed; the analytic functor is what the planner evaluates), so it gets its own documented tolerance.
OUTLINE_TOL = {"sdCutDisk": 0.45}
AREA_TOL = {"sdCutDisk": 0.17}


def test_registry_keys(oracle_mod):
    ids = [oracle_mod.lib().orc_shape_id(n.encode()) for n in ANALYTIC]
    assert ids == list(range(16))
    assert oracle_mod.lib().orc_shape_id(b"no_such_shape") == 17  # Polygon fallback (sw_manager.hpp:363-372)
    assert oracle_mod.lib().orc_shape_id(b"anyMesh") == 17


def test_outline_vertices_of_reference_meshes_lie_on_zero_level_set(oracle_mod, outlines):
    assert set(outlines) == {"sdArc", "sdCutDisk", "sdHeart", "sdHorseshoe", "sdOrientedVesica", "sdPie", "sdPie2",
                             "sdRhombus", "sdRoundedCross", "sdRoundedX", "sdTunnel", "sdUnevenCapsule", "star"}
    for name, rec in outlines.items():
        s = oracle_mod.shape_sdf(name, _rel(rec["outline_xy"]))
        tol = OUTLINE_TOL.get(name, 0.09)
        assert np.abs(s).max() <= tol, (name, np.abs(s).max())
        # every mesh vertex is inside or on the analytic shape
        s_all = oracle_mod.shape_sdf(name, _rel(rec["all_top_xy"]))
        assert s_all.max() <= tol, (name, s_all.max())


def test_area_enclosed_matches_reference_meshes(oracle_mod, outlines):
    g = np.linspace(-8.0, 8.0, 1601)
    X, Y = np.meshgrid(g, g)
    R = np.c_[X.ravel(), Y.ravel(), np.zeros(X.size)]
    cell = (g[1] - g[0]) ** 2
    for name, rec in outlines.items():
        area = (oracle_mod.shape_sdf(name, R) < 0).sum() * cell
        ratio = rec["top_area"] / area
        assert abs(ratio - 1.0) <= AREA_TOL.get(name, 0.085), (name, ratio)


def test_known_answers(oracle_mod):
    f = lambda n, x, y: float(oracle_mod.shape_sdf(n, _rel([[x, y]]))[0])
    # star: r = 2.8 tip on +y; outside the tip the nearest feature is the tip vertex
    assert abs(f("star", 0.0, 2.8)) < 1e-12
    assert abs(f("star", 0.0, 3.8) - 1.0) < 1e-12
    assert abs(f("star", 0.0, 10.0) - 7.2) < 1e-12
    assert f("star", 0.0, 0.0) < 0
    # five-fold symmetry of the star
    for k in range(1, 5):
        a = 2 * math.pi * k / 5
        x, y = -math.sin(a) * 3.3, math.cos(a) * 3.3
        assert abs(f("star", x, y) - 0.5) < 1e-9
    # rhombus b = (1, 4.5)
    assert abs(f("sdRhombus", 1.0, 0.0)) < 1e-12 and abs(f("sdRhombus", 0.0, 4.5)) < 1e-12
    assert abs(f("sdRhombus", 0.0, 0.0) + 4.5 / math.hypot(1.0, 4.5)) < 1e-12
    # trapezoid r1 = 1 (y = -2), r2 = 3 (y = +2)
    assert abs(f("sdTrapezoid", 3.0, 2.0)) < 1e-12 and abs(f("sdTrapezoid", 1.0, -2.0)) < 1e-12
    assert abs(f("sdTrapezoid", 0.0, 0.0) + 8.0 / math.sqrt(20.0)) < 1e-12
    assert abs(f("sdTrapezoid", 0.0, 3.0) - 1.0) < 1e-12
    # tunnel wh = (2.5, 1.5): the functor flips y, so the box (depth 1.5) is on -y and the half disc (radius 2.5) on +y
    assert abs(f("sdTunnel", 0.0, 0.0) + 1.5) < 1e-12
    assert abs(f("sdTunnel", 0.0, -1.5)) < 1e-12
    assert abs(f("sdTunnel", 0.0, 2.5)) < 1e-12
    assert abs(f("sdTunnel", 2.5, -0.5)) < 1e-12
    # cut disk r = 5, h = 2 (kept part: y > h)
    assert abs(f("sdCutDisk", 0.0, 0.0) - 2.0) < 1e-12
    assert abs(f("sdCutDisk", 0.0, 4.0) + 1.0) < 1e-12
    assert abs(f("sdCutDisk", 0.0, 6.0) - 1.0) < 1e-12
    # uneven capsule r1 = 2, r2 = 1, h = 5
    assert abs(f("sdUnevenCapsule", 0.0, -2.0)) < 1e-12 and abs(f("sdUnevenCapsule", 0.0, 6.0)) < 1e-12
    assert abs(f("sdUnevenCapsule", 0.0, 0.0) + 2.0) < 1e-12
    # pie: apex on the boundary, far along +y-ish outside
    assert abs(f("sdPie", 0.0, 0.0)) < 1e-12 and abs(f("sdPie2", 0.0, 0.0)) < 1e-12
    # arc: ra = 2.3333, rb = 0.5, centred on +y
    assert abs(f("sdArc", 0.0, 2.3333) + 0.5) < 1e-12 and abs(f("sdArc", 0.0, 2.8333)) < 1e-12
    # rounded X: w = 3 (5 for bigX), r = 0.25; far on the diagonal the nearest feature is the segment end (w/2, w/2)
    assert abs(f("sdRoundedX", 3.0, 3.0) - (math.hypot(1.5, 1.5) - 0.25)) < 1e-12
    assert abs(f("bigX", 4.0, 4.0) - (math.hypot(1.5, 1.5) - 0.25)) < 1e-12
    assert abs(f("sdRoundedX", 0.0, 0.0) + 0.25) < 1e-12
    # heart (scaled by 4): bottom tip at the origin
    assert abs(f("sdHeart", 0.0, 0.0)) < 1e-12
    # moon d = 0.8, ra = 3, rb = 2.4: outer circle on -x side
    assert abs(f("sdMoon", -3.0, 0.0)) < 1e-12 and abs(f("sdMoon", -4.0, 0.0) - 1.0) < 1e-12
    # circle radius 1
    assert abs(f("Circle", 3.0, 4.0) - 4.0) < 1e-12


def test_eikonal_property(oracle_mod):
    rng = np.random.default_rng(7)
    pts = _rel(rng.uniform(-7.0, 7.0, size=(4000, 2)))
    for name in ANALYTIC + ["Circle"]:
        g = oracle_mod.shape_grad1(name, pts)
        n = np.hypot(g[:, 0], g[:, 1])
        assert np.all(g[:, 2] == 0.0)
        # exact SDFs: |grad| = 1 away from the medial axis (FD with dx = 1e-6 straddles it for a few samples)
        assert abs(np.median(n) - 1.0) < 1e-6, (name, np.median(n))
        assert np.mean(np.abs(n - 1.0) < 1e-4) > 0.97, (name, np.mean(np.abs(n - 1.0) < 1e-4))


def test_fd_gradient_follows_reference_macro(oracle_mod):
    # Shape.hpp:35-53: g = (f((x-dx)+2dx, y) - f(x-dx, y), f(x, (y-dx)+2dx) - f(x, y-dx), 0) / (2 dx)
    rng = np.random.default_rng(11)
    p = rng.uniform(-5, 5, size=(50, 2))
    dx = 0.000001
    for name in ("star", "sdHorseshoe", "sdMoon"):
        g = oracle_mod.shape_grad1(name, _rel(p))
        xm = p[:, 0] - dx
        xp = xm + 2 * dx
        ym = p[:, 1] - dx
        yp = ym + 2 * dx
        fx = oracle_mod.shape_sdf(name, _rel(np.c_[xp, p[:, 1]])) - oracle_mod.shape_sdf(name, _rel(np.c_[xm, p[:, 1]]))
        fy = oracle_mod.shape_sdf(name, _rel(np.c_[p[:, 0], yp])) - oracle_mod.shape_sdf(name, _rel(np.c_[p[:, 0], ym]))
        assert np.array_equal(g[:, 0], fx / (2 * dx)) and np.array_equal(g[:, 1], fy / (2 * dx))


def test_body_frame_pretransform(oracle_mod):
    # Shape.hpp:281-294: P = ((rel - trans) * Rotate).head(2), Rotate = Rz(poly_params[2] deg), row-vector times matrix
    rng = np.random.default_rng(3)
    rel = _rel(rng.uniform(-6, 6, size=(200, 2)))
    tx, ty, deg = 0.7, -0.4, 33.0
    yaw = deg * 3.14159265358979323846 / 180.0
    R = np.array([[math.cos(yaw), -math.sin(yaw)], [math.sin(yaw), math.cos(yaw)]])
    P = (rel[:, :2] - np.array([tx, ty])) @ R
    for name in ("star", "sdHorseshoe", "sdTunnel"):
        a = oracle_mod.shape_sdf(name, rel, poly_params=(tx, ty, deg))
        b = oracle_mod.shape_sdf(name, _rel(P))
        assert np.abs(a - b).max() < 1e-12


def test_polygon_fallback(oracle_mod):
    # sw_manager.hpp:363-372: unknown names -> rectangle (6,-0.1),(6,0.1),(-6,0.1),(-6,-0.1)
    f = lambda x, y: float(oracle_mod.shape_sdf("mesh_shape", _rel([[x, y]]))[0])
    assert abs(f(0.0, 0.0) + 0.1) < 1e-12
    assert abs(f(0.0, 1.1) - 1.0) < 1e-12
    assert abs(f(7.0, 0.0) - 1.0) < 1e-12
    assert abs(f(9.0, 4.1) - 5.0) < 1e-12  # corner (6, 0.1)
    g = oracle_mod.shape_grad1("mesh_shape", _rel([[0.0, 1.1], [7.0, 0.0], [0.0, 0.05]]))
    assert np.allclose(g[0], [0, 1, 0]) and np.allclose(g[1], [1, 0, 0])
    assert np.allclose(g[2], [0, 1, 0])  # inside: -(p - c) normalised, c on the top edge
    # custom polygon (triangle)
    tri = [0, 0, 4, 0, 0, 3]
    s = oracle_mod.shape_sdf("custom", _rel([[1, 1], [5, 0], [-1, -1]]), polygon=tri)
    assert s[0] < 0 and abs(s[1] - 1.0) < 1e-12 and abs(s[2] - math.sqrt(2)) < 1e-12


def test_circumradius_bound(oracle_mod):
    """The kernels skip lattice samples of choiceTInit's first layer whose pose is farther from the query point than
    (current minimum + R): that is exact iff sdf(q) >= |q| - R for every q.  Check the library's R (svsdf_shape_bound_radius,
    host code) against the oracle's functors: far field, near field, with and without a body-frame pre-transform, and the
    Polygon fallback with a custom outline.  Also: R is not loose (a point of the shape lies within 0.12 of the circle)."""
    from implicit_svsdf_planner_b200 import api

    rng = np.random.default_rng(12)
    far = np.c_[rng.uniform(-80, 80, (200_000, 2)), np.zeros(200_000)]
    rr, th = rng.uniform(0, 14, 200_000), rng.uniform(0, 2 * np.pi, 200_000)
    near = np.c_[rr * np.cos(th), rr * np.sin(th), np.zeros(200_000)]
    pts = np.r_[far, near]
    d = np.linalg.norm(pts[:, :2], axis=1)
    cases = [(n, (0.0, 0.0, 0.0), None) for n in ANALYTIC + ["Circle", "no_such_shape"]]
    cases += [("star", (0.6, -0.3, 25.0), None), ("sdHorseshoe", (-1.0, 2.0, -70.0), None),
              ("custom", (0.0, 0.0, 0.0), [3.0, -1.0, 3.0, 1.0, 0.0, 2.5, -3.0, 1.0, -3.0, -1.0])]
    for name, pp, poly in cases:
        R = api.shape_bound_radius(name, pp, poly)
        f = oracle_mod.shape_sdf(name, pts, poly_params=pp, polygon=poly)
        slack = f - (d - R)
        assert slack.min() >= 0.03, (name, pp, R, slack.min())
        ring = np.c_[(R - 0.12) * np.cos(np.linspace(0, 2 * np.pi, 20001)), (R - 0.12) * np.sin(np.linspace(0, 2 * np.pi, 20001)), np.zeros(20001)]
        if pp == (0.0, 0.0, 0.0):
            assert oracle_mod.shape_sdf(name, ring, poly_params=pp, polygon=poly).min() <= 0.0, (name, R)
