"""GPU (-m gpu): K3, device-side query-point construction from the byte-packed map kernel (csrc/svsdf_extract.cu), against
the numpy restatement of plan_manager.cpp:156-175 / PCSmap_manager.h:184-219 (batch.extract_query_points).  Integer / byte
work: results must be identical (same cells, same order, bit-identical cell-centre coordinates)."""
import numpy as np
import pytest

from implicit_svsdf_planner_b200 import api, batch, scenes

pytestmark = pytest.mark.gpu


def _check(gm, wps, half, kernel_size=17, keepout=None, clearance=0.0):
    ctx = api.Context("star")
    X, Y = gm.shape
    ctx.set_map(batch.pack_map_kernel(gm.occ, kernel_size), X, Y, kernel_size, gm.origin, gm.res)
    n = ctx.extract_points(wps, half, keepout, clearance)
    got = ctx.get_points()
    ref = batch.extract_query_points(gm, wps, half, keepout, clearance)[:, :2]
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
