"""CPU tests of the oracle's triangle-mesh functor (BasicShape::getonlySDF_igl, Shape.hpp:332-340) and of the host-side
.obj reader.  Pins: closed-form answers on a cube, the 2-D polygon SDF of the extruded outline, and the reference's OWN
fast-winding-number code (compiled from /root/reference into oracle/_ref; outputs committed in tests/golden/fwn_ref.npz)."""
import os

import numpy as np
import pytest

from implicit_svsdf_planner_b200 import api, scenes

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FWN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_fwn.so")


def cube(h=1.0):
    V = np.array([[x, y, z] for x in (-h, h) for y in (-h, h) for z in (-h, h)], dtype=np.float64)
    # outward-oriented faces of the cube (vertex index = 4*ix + 2*iy + iz)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    F = []
    for a, b, c, d in quads:
        F += [[a, b, c], [a, c, d]]
    return V, np.asarray(F, dtype=np.int32)


def test_cube_known_answers(oracle_mod):
    m = cube(1.0)
    Q = np.array([[0, 0, 0], [0.5, 0.2, 0], [2.0, 0, 0], [2.0, 2.0, 0], [0, -3.0, 0], [2.0, 2.0, 2.0], [0.25, 0, 0.9]], dtype=np.float64)
    w = oracle_mod.mesh_eval(m, Q, "winding")
    assert np.allclose(w, [1, 1, 0, 0, 0, 0, 1], atol=1e-14)
    d2 = oracle_mod.mesh_eval(m, Q, "sqr_distance")
    assert np.allclose(d2, [1.0, 0.25, 1.0, 2.0, 4.0, 3.0, 0.01], atol=1e-14)
    sdf = oracle_mod.mesh_eval(m, Q, "sdf")
    assert np.allclose(sdf, [-1.0, -0.5, 1.0, np.sqrt(2.0), 2.0, np.sqrt(3.0), -0.1], atol=1e-13)
    # orientation matters: the inward-oriented cube has winding number -1 inside -> (1 - 2w) = 3 (the reference's formula)
    m_in = (m[0], m[1][:, ::-1].copy())
    assert np.allclose(oracle_mod.mesh_eval(m_in, Q[:1], "winding"), [-1.0], atol=1e-14)


def test_extruded_outline_matches_the_polygon_sdf_in_the_mid_plane(oracle_mod):
    outline = scenes.star_outline(n_per_edge=3)
    m = scenes.extrude_outline(outline, half_height=0.49)
    rng = np.random.default_rng(5)
    Q = np.c_[rng.uniform(-4, 4, size=(4000, 2)), np.zeros(4000)]
    sd = oracle_mod.mesh_eval(m, Q, "sdf")
    poly = oracle_mod.shape_sdf("custom_poly", Q, polygon=outline)
    out = poly > 0
    assert np.abs(sd - poly)[out].max() < 1e-13           # outside: the in-plane distance
    assert (np.sign(sd) == np.sign(poly)).all()
    assert np.allclose(sd[~out], np.maximum(poly[~out], -0.49), atol=1e-13)  # inside: saturates at the caps
    w = oracle_mod.mesh_eval(m, Q, "winding")
    assert np.abs(w - np.round(w)).max() < 1e-13          # closed mesh: integer winding number
    # FD gradient (Shape.hpp:35-53): unit length away from creases, zero z component
    g = oracle_mod.mesh_eval(m, Q[out][:500], "grad1")
    nrm = np.linalg.norm(g[:, :2], axis=1)
    assert (np.abs(nrm - 1.0) < 1e-6).mean() > 0.98 and np.all(g[:, 2] == 0.0)


def test_poly_params_move_the_mesh_vertices(oracle_mod):
    """BasicShape's constructor applies poly_params to the vertices (R v + trans, Shape.hpp:285-302): the level set moves
    WITH the transform (unlike the analytic functors, which transform the query)."""
    outline = scenes.star_outline()
    m = scenes.extrude_outline(outline)
    pp = (0.6, -0.3, 25.0)
    th = np.deg2rad(pp[2])
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    rng = np.random.default_rng(6)
    Q = np.c_[rng.uniform(-4, 4, size=(500, 2)), np.zeros(500)]
    Qb = Q.copy()
    Qb[:, :2] = (Q[:, :2] - np.array(pp[:2])) @ R  # R^T (q - trans)
    a = oracle_mod.mesh_eval(m, Q, "sdf", poly_params=pp)
    b = oracle_mod.mesh_eval(m, Qb, "sdf")
    assert np.abs(a - b).max() < 1e-12


@pytest.mark.parametrize("name", ["star", "sdHorseshoe"])
def test_exact_winding_number_against_the_reference_fwn_golden(oracle_mod, name):
    """w_ref was produced by the reference's own igl/HDK code on its own shapes/*.obj (tests/golden/make_fwn_golden.py).
    It is a float, order-2 Barnes-Hut approximation of the sum the oracle evaluates exactly; the difference is the
    approximation error of the reference (2.1e-3 at worst), and rounding w_ref gives the oracle's integer."""
    g = np.load(os.path.join(HERE, "golden", "fwn_ref.npz"))
    m = (g[name + "_V"], g[name + "_F"])
    Q, w_ref = g[name + "_Q"], g[name + "_w_ref"]
    w = oracle_mod.mesh_eval(m, Q, "winding")
    assert np.abs(w - np.round(w)).max() < 1e-13
    d2 = oracle_mod.mesh_eval(m, Q, "sqr_distance")
    far = d2 > 1e-4  # on the surface the winding number jumps; float vs double vertices decide the side
    assert np.abs(w - w_ref)[far].max() < 5e-3, np.abs(w - w_ref)[far].max()
    assert np.array_equal(np.round(w_ref[far]), np.round(w[far]))
    # hence getonlySDF_igl = (1 - 2 w) sqrt(d2) of the reference is within 1 % of the exact signed distance
    sdf = oracle_mod.mesh_eval(m, Q, "sdf")
    sdf_ref = (1.0 - 2.0 * w_ref) * np.sqrt(d2)
    assert np.abs(sdf - sdf_ref)[far].max() <= 1.0e-2 * np.abs(sdf[far]).max()


@pytest.mark.skipif(not os.path.exists(REF_FWN), reason="oracle/_ref/libref_fwn.so not built (needs /root/reference)")
def test_reference_fwn_live_on_a_synthetic_mesh(oracle_mod):
    import sys

    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_fwn_golden as mk

    m = scenes.extrude_outline(scenes.star_outline(n_per_edge=4))
    rng = np.random.default_rng(9)
    Q = np.c_[rng.uniform(-4, 4, size=(3000, 2)), np.zeros(3000)]
    w_ref = mk.ref_fwn(m[0], m[1], Q)
    w = oracle_mod.mesh_eval(m, Q, "winding")
    far = oracle_mod.mesh_eval(m, Q, "sqr_distance") > 1e-4
    assert np.abs(w - w_ref)[far].max() < 5e-3


def test_obj_reader_host(tmp_path):
    p = tmp_path / "m.obj"
    p.write_text("# comment\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\n"
                 "vn 0 0 1\nf 1 2 3 4\nf 1/1/1 2/2/1 5/3/1\nf -4//1 -3//1 -1//1\n")
    V, F = api.read_obj(str(p))
    V2, F2 = scenes.load_obj(str(p))
    assert np.array_equal(V, V2) and np.array_equal(F, F2)
    assert V.shape == (5, 3) and F.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [1, 2, 4]]
    with pytest.raises(api.SvsdfError):
        api.read_obj(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(api.SvsdfError):
        api.read_obj(str(bad))
