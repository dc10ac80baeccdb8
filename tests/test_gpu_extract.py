"""GPU (-m gpu): K3, device-side query-point construction from the byte-packed map kernel (csrc/svsdf_extract.cu), against
the CPU restatement of plan_manager.cpp:131-175 / PCSmap_manager.h:184-219 / Gridmap3D.cpp (oracle/k3_points.py).  Integer / byte
work: results must be identical (same cells, same order, bit-identical cell-centre coordinates) — flat maps and the reference's
3-D maps (z-stacked voxels give repeated (x, y) points; the first waypoint skips the box around tmp_pos = (999, 999, 999))."""
import os
import numpy as np
import pytest

from implicit_svsdf_planner_b200 import api, batch, scenes
from oracle import k3_points

pytestmark = pytest.mark.gpu


def _check(gm, wps, half, kernel_size=17, keepout=None, clearance=0.0):
    ctx = api.Context("star")
    X, Y = gm.shape
    ctx.set_map(batch.pack_map_kernel(gm.occ, kernel_size), X, Y, kernel_size, gm.origin, gm.res)
    n = ctx.extract_points(wps, half, keepout, clearance)
    got = ctx.get_points()
    ref = k3_points.query_points_2d(gm.occ, gm.origin, gm.res, wps, half, keepout, clearance)[:, :2]
    assert n == ref.shape[0] == got.shape[0], (n, ref.shape, got.shape)
    assert np.array_equal(got, ref)
    return ctx, ref


def test_extraction_matches_numpy_restatement_on_random_maps():
    rng = np.random.default_rng(0)
    for (X, Y, res, dens, ks) in ((40, 40, 1.0, 1.0, 17), (97, 61, 0.5, 0.35, 17), (130, 333, 0.25, 0.2, 9), (257, 64, 0.1, 0.5, 3)):
        occ = rng.random((X, Y)) < dens
        gm = batch.GridMap(occ=occ, origin=np.array([-3.7, 2.25]), res=res)
        ext = np.array([X, Y]) * res
        for trial in range(4):
            W = int(rng.integers(1, 9))
            wps = gm.origin + rng.uniform(-0.1, 1.1, size=(W, 2)) * ext  # some boxes stick out of the map (projInMap)
            half = float(rng.uniform(0.6, 9.0))
            _check(gm, wps, half, ks)


def test_out_of_last_box_rule_and_dedup():
    occ = np.ones((40, 40), dtype=bool)
    gm = batch.GridMap(occ=occ, origin=np.zeros(2), res=1.0)
    ctx, ref = _check(gm, np.array([[10.5, 10.5]]), 3.0)
    assert ref.shape[0] == 49
    ctx, ref = _check(gm, np.array([[10.5, 10.5], [12.5, 10.5]]), 3.0)
    assert ref.shape[0] == 49 + 14
    # identical consecutive waypoints add nothing; a cell skipped because of the previous box is still found via an earlier box
    ctx, ref = _check(gm, np.array([[10.5, 10.5], [10.5, 10.5], [12.5, 10.5], [10.5, 10.5]]), 3.0)
    assert np.unique(ref, axis=0).shape[0] == ref.shape[0]
    # empty map, empty result
    gm0 = batch.GridMap(occ=np.zeros((33, 35), dtype=bool), origin=np.zeros(2), res=0.5)
    ctx, ref = _check(gm0, np.array([[5.0, 5.0]]), 2.0)
    assert ref.shape[0] == 0


def test_keepout_option_and_use_as_query_set(oracle_mod):
    gm = batch.make_random_map(extent=60.0, res=0.25, density=0.25, seed=5)
    start, goal = scenes.START_GOAL["star"]
    init_s, final_s, q, T = scenes.make_trajectory("star", 8, 77, start, goal)
    b = scenes.minco_dense(init_s, final_s, q, T)
    wps = np.concatenate([init_s[:2, :1], q[:2], final_s[:2, :1]], axis=1).T
    half = 17 / 3.0
    ko = batch.keepout_samples(b, T)
    ctx, ref = _check(gm, wps, half, 17, ko, 2.9)
    assert 1000 < ref.shape[0] < 20000
    d = np.sqrt(((ref[:, None, :] - ko[None, :, :]) ** 2).sum(-1)).min(axis=1)
    assert d.min() > 2.9
    # the extracted set is the context's resident query set: cost+gradient equals the one from host-uploaded points
    co = np.ascontiguousarray(b.T).reshape(-1)
    c1, gT1, gC1 = ctx.cost_grad(T, co)
    ctx2 = api.Context("star")
    ctx2.set_points(np.c_[ref, np.zeros(len(ref))])
    c2, gT2, gC2 = ctx2.cost_grad(T, co)
    assert c1 == c2 and np.array_equal(gC1, gC2) and np.array_equal(gT1, gT2)
    orc = oracle_mod.Oracle("star", threads=oracle_mod.num_procs())
    orc.set_points(np.c_[ref, np.zeros(len(ref))])
    c0, gT0, gC0, _, _ = orc.cost_grad(T, co)
    assert abs(c1 - c0) <= 1e-12 * abs(c0)


def _check3d(gm, wps, half3, kernel_size=17):
    X, Y, Z = gm.size
    ctx = api.Context("star")
    ctx.set_map3d(k3_points.generate_map_kernel(gm, kernel_size), X, Y, Z, kernel_size, gm.boundary_min, gm.res)
    n = ctx.extract_points3d(wps, half3)
    got = ctx.get_points()
    ref = k3_points.query_points(gm, wps, half3)  # sorted by unifiedID = k * X * Y + j * X + i
    # the device emits layer by layer, ascending (i * Y + j) within a layer
    idx = np.array([gm.index_of_center(p) for p in ref]).reshape(-1, 3)
    order = np.lexsort((idx[:, 1], idx[:, 0], idx[:, 2])) if len(ref) else np.zeros(0, dtype=int)
    assert n == len(ref) == len(got), (n, len(ref), len(got))
    assert np.array_equal(got, ref[order][:, :2])
    return ctx, ref


def test_3d_map_of_the_reference_star_scene():
    """pcds/map_star.pcd (tests/golden/map_star_pcd.npz): 31 x 76 x 9 voxels, 148 occupied, obstacles stacked in z."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "map_star_pcd.npz"))
    gm = k3_points.gridmap_from_cloud(g["points"], float(g["occupancy_resolution"]), int(g["sta_threshold"]))
    res, ks = float(g["occupancy_resolution"]), int(g["kernel_size"])
    n = int(np.ceil(np.linalg.norm(g["end"] - g["start"]) / res)) + 1
    path = g["start"][None, :] + np.linspace(0.0, 1.0, n)[:, None] * (g["end"] - g["start"])[None, :]
    wps = k3_points.waypoints_of_path(path, float(g["traj_parlength"]), res)
    ctx, ref = _check3d(gm, wps, [ks * res / 3.0] * 3, ks)
    assert len(ref) > 20
    # one box over everything: all 148 voxels but the far corner one — occupied in this scene, and skipped for the first waypoint
    # because the box around tmp_pos = (999, 999, 999) projects onto it
    assert gm.occ[-1, -1, -1]
    ctx, ref = _check3d(gm, np.array([[15.0, 37.0, 4.0]]), [100.0] * 3, ks)
    assert len(ref) == 147
    # layer 0 of the 3-D upload is the 2-D map the front end sees: the flat API on that layer gives layer 0's cells (minus ITS far
    # corner cell (30, 75), which a one-layer map skips for the same reason)
    ctx2 = api.Context("star")
    ctx2.set_map(k3_points.generate_map_kernel_2d(gm, ks), gm.size[0], gm.size[1], ks, gm.boundary_min[:2], res)
    n2 = ctx2.extract_points(np.array([[15.0, 37.0]]), 100.0)
    flat = k3_points.query_points_2d(gm.occ[:, :, 0], gm.boundary_min[:2], res, np.array([[15.0, 37.0]]), 100.0)
    assert n2 == len(flat) == 140 - int(gm.occ[-1, -1, 0]) and np.array_equal(ctx2.get_points(), flat[:, :2])
    layer0 = ctx.get_points()[:140]
    assert {tuple(p) for p in ctx2.get_points()} <= {tuple(p) for p in layer0}


def test_3d_extraction_on_random_maps_and_the_tmp_pos_rule():
    rng = np.random.default_rng(12)
    for (X, Y, Z, res, dens) in ((20, 17, 4, 1.0, 1.0), (45, 33, 7, 0.5, 0.3), (70, 19, 2, 0.25, 0.5)):
        occ = rng.random((X, Y, Z)) < dens
        lo = np.array([-2.5, 1.25, 0.0])
        gm = k3_points.GridMap3D(boundary_min=lo, boundary_max=lo + np.array([X, Y, Z]) * res, res=res, occ=occ)
        ext = np.array([X, Y, Z]) * res
        for trial in range(3):
            W = int(rng.integers(1, 7))
            wps = lo + rng.uniform(-0.1, 1.1, size=(W, 3)) * ext
            half3 = rng.uniform(0.4, 0.6 * ext.max(), size=3)
            _check3d(gm, wps, half3, 9)
    # full map, one all-covering box: everything but the far corner voxel (the box around tmp_pos = (999, 999, 999))
    occ = np.ones((6, 5, 2), dtype=bool)
    gm = k3_points.GridMap3D(boundary_min=np.zeros(3), boundary_max=np.array([6.0, 5.0, 2.0]), res=1.0, occ=occ)
    ctx, ref = _check3d(gm, np.array([[3.0, 2.5, 1.0]]), [10.0] * 3, 5)
    assert len(ref) == 59
    # cost over stacked voxels: a column with two occupied layers counts twice
    pts = ctx.get_points()
    assert len(np.unique(pts, axis=0)) == 30 and len(pts) == 59


def test_extract_errors():
    ctx = api.Context("star")
    with pytest.raises(api.SvsdfError):  # map not set
        ctx.extract_points(np.zeros((1, 2)), 1.0)
    with pytest.raises(api.SvsdfError):  # even kernel size
        ctx.set_map(np.zeros((10, 2), dtype=np.uint8), 4, 4, 6, (0, 0), 1.0)


def test_map_survives_growing_queries(scene2k):
    """Regression: growing the per-point query buffers (svsdf_query with more points than before) must not touch the map."""
    rng = np.random.default_rng(5)
    occ = rng.random((60, 60)) < 0.3
    gm = batch.GridMap(occ=occ, origin=np.zeros(2), res=1.0)
    ctx, ref = _check(gm, np.array([[20.5, 20.5], [30.0, 31.0]]), 4.0)
    sc = scene2k
    co = sc.coeffs_colmajor()
    p = np.c_[sc.points[:, :2], np.zeros(sc.P)]
    ctx.query(sc.T, co, p[:10])
    ctx.query(sc.T, co, p)              # larger than any query before: buffers are re-allocated
    ctx.query(sc.T, co, np.r_[p, p, p])  # and again
    n = ctx.extract_points(np.array([[20.5, 20.5], [30.0, 31.0]]), 4.0)
    assert n == ref.shape[0] and np.array_equal(ctx.get_points(), ref)
    ctx.close()
