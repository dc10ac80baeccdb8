"""CPU tests of oracle/k3_points.py — the restatement of the reference's point-cloud -> occupancy map -> query points chain
(PCSmap_manager.cpp:113-190, Gridmap3D.cpp, PCSmap_manager.h:118-219, plan_manager.cpp:131-175) that the device extraction kernels are
checked against.  Pin: the reference's own scene pcds/map_star.pcd gives 148 occupied voxels (tests/golden/map_star_pcd.npz)."""
import os

import numpy as np

from implicit_svsdf_planner_b200 import batch
from oracle import k3_points as k3

HERE = os.path.dirname(os.path.abspath(__file__))


def star_map():
    g = np.load(os.path.join(HERE, "golden", "map_star_pcd.npz"))
    return g, k3.gridmap_from_cloud(g["points"], float(g["occupancy_resolution"]), int(g["sta_threshold"]))


def test_map_star_pcd_gives_148_voxels():
    g, gm = star_map()
    assert gm.size == (31, 76, 9)  # ceil(extent / 1.0) per axis (Gridmap3D.cpp:29-31)
    assert int(gm.occ.sum()) == int(g["occupied_voxels"]) == 148
    assert gm.occ.sum(axis=(0, 1)).tolist() == [140, 1, 1, 0, 0, 2, 2, 1, 1]  # the scene has obstacles stacked in z
    # every cloud point falls into an occupied voxel; the boundary points land in the first / last cells
    for p in g["points"].astype(np.float64):
        assert gm.occ[gm.grid_index(p)]
    assert gm.grid_index(gm.boundary_max) == (30, 75, 8) and gm.grid_index(gm.boundary_min) == (0, 0, 0)
    assert gm.grid_index(gm.boundary_max + 1e-9) == (0, 0, 0)  # isInMap fails -> (0, 0, 0), Gridmap3D.cpp:139-142


def test_query_points_of_the_star_scene():
    """A straight front-end path from the scene's start to its end, cut into waypoints as generateTraj does."""
    g, gm = star_map()
    res, ks = float(g["occupancy_resolution"]), int(g["kernel_size"])
    n = int(np.ceil(np.linalg.norm(g["end"] - g["start"]) / res)) + 1
    path = g["start"][None, :] + np.linspace(0.0, 1.0, n)[:, None] * (g["end"] - g["start"])[None, :]
    wps = k3.waypoints_of_path(path, float(g["traj_parlength"]), res)
    assert len(wps) == (n - 2) // 3 and np.allclose(wps[0], path[3])  # index_gap = ceil(3.0 / 1.0)
    half = ks * res / 3.0
    pts = k3.query_points(gm, wps, [half] * 3)
    ids = {gm.unified_id(*gm.index_of_center(p)) for p in pts}
    assert len(ids) == len(pts) > 20                                   # de-duplicated by voxel id
    assert all(gm.occ[gm.index_of_center(p)] for p in pts)                  # only occupied voxels
    # brute force: a voxel is a query point iff it is occupied and inside some waypoint box but not only through boxes whose
    # predecessor box also contains it
    want = set()
    last = k3.TMP_POS
    for wp in wps:
        c1, c2 = gm.grid_index(gm.proj_in_map(wp - half)), gm.grid_index(gm.proj_in_map(wp + half))
        l1, l2 = gm.grid_index(gm.proj_in_map(last - half)), gm.grid_index(gm.proj_in_map(last + half))
        for i, j, k in zip(*np.nonzero(gm.occ)):
            inb = all(c1[a] <= v <= c2[a] for a, v in enumerate((i, j, k)))
            inl = all(l1[a] <= v <= l2[a] for a, v in enumerate((i, j, k)))
            if inb and not inl:
                want.add(gm.unified_id(i, j, k))
        last = wp
    assert ids == want
    # z-stacked voxels: several query points share (x, y)
    xy = np.round(pts[:, :2], 9)
    assert len(np.unique(xy, axis=0)) <= len(pts)


def test_first_waypoint_skips_the_box_around_tmp_pos():
    occ = np.ones((6, 5, 2), dtype=bool)
    gm = k3.GridMap3D(boundary_min=np.zeros(3), boundary_max=np.array([6.0, 5.0, 2.0]), res=1.0, occ=occ)
    pts = k3.query_points(gm, np.array([[3.0, 2.5, 1.0]]), [10.0] * 3)   # one box over the whole map
    assert len(pts) == 6 * 5 * 2 - 1                                      # all but the far corner voxel (5, 4, 1)
    assert not any(np.allclose(p, [5.5, 4.5, 1.5]) for p in pts)
    pts2 = k3.query_points(gm, np.array([[3.0, 2.5, 1.0], [3.0, 2.5, 1.0]]), [10.0] * 3)
    assert len(pts2) == len(pts)                                          # the second box equals its predecessor: nothing new
    pts3 = k3.query_points(gm, np.array([[0.5, 0.5, 0.5], [3.0, 2.5, 1.0]]), [0.6, 0.6, 0.6])
    assert len(pts3) > 8 and any(np.allclose(p, [0.5, 0.5, 0.5]) for p in pts3)


def test_packed_kernels_and_the_flat_case_agree_with_the_product_packing():
    rng = np.random.default_rng(3)
    occ = rng.random((13, 21, 5)) < 0.3
    gm = k3.GridMap3D(boundary_min=np.array([-1.0, 2.0, 0.0]), boundary_max=np.array([-1.0, 2.0, 0.0]) + np.array([13, 21, 5]) * 0.5, res=0.5, occ=occ)
    k2 = k3.generate_map_kernel_2d(gm, 17)
    assert np.array_equal(k2, batch.pack_map_kernel(occ[:, :, 0], 17))   # the product's packer (host code of the batch mode)
    kk = k3.generate_map_kernel(gm, 17)
    assert kk.shape == (13 + 16, 21 + 16, (5 + 16 + 7) // 8) and int(np.unpackbits(kk).sum()) == int(occ.sum())
    # flat wrapper == 3-D restatement on a one-layer map
    wps = np.array([[0.3, 4.1], [2.2, 6.0], [2.4, 6.1]])
    flat = k3.query_points_2d(occ[:, :, 0], gm.boundary_min[:2], 0.5, wps, 1.2)
    gm1 = k3.gridmap_2d(occ[:, :, 0], gm.boundary_min[:2], 0.5)
    full = k3.query_points(gm1, np.c_[wps, np.full(3, 0.25)], [1.2] * 3)
    assert len(flat) == len(full) and {tuple(np.round(p[:2], 9)) for p in flat} == {tuple(np.round(p[:2], 9)) for p in full}


def test_product_host_glue_of_the_plan_chain_equals_the_restatement():
    """implicit_svsdf_planner_b200/plan.py (the product's numpy glue for one whole plan): map from cloud, 3-D packed map, waypoint rule."""
    from implicit_svsdf_planner_b200 import plan

    g, gm = star_map()
    cm = plan.gridmap3d_from_cloud(g["points"], float(g["occupancy_resolution"]), int(g["sta_threshold"]))
    assert np.array_equal(cm.occ, gm.occ) and np.array_equal(cm.boundary_min, gm.boundary_min) and np.array_equal(cm.boundary_max, gm.boundary_max)
    assert np.array_equal(plan.pack_map_kernel3d(cm.occ, 17), k3.generate_map_kernel(gm, 17))
    rng = np.random.default_rng(4)
    for n in (3, 4, 7, 20, 64):
        path = rng.uniform(0, 30, size=(n, 3))
        for L, res in ((3.0, 1.0), (3.0, 0.25), (0.7, 1.0)):
            idx, w = plan.waypoints_of_path(path, L, res)
            assert np.array_equal(w, k3.waypoints_of_path(path, L, res)) and np.array_equal(path[idx], w)
    # sta_threshold > 1: only voxels hit by several points stay
    cm2 = plan.gridmap3d_from_cloud(g["points"], 1.0, 2)
    gm2 = k3.gridmap_from_cloud(g["points"], 1.0, 2)
    assert np.array_equal(cm2.occ, gm2.occ) and cm2.occ.sum() < cm.occ.sum()
