"""GPU (-m gpu): one whole plan on the reference's own star scene (pcds/map_star.pcd, trajectory_star.txt, config/star.yaml; fixture
tests/golden/map_star_pcd.npz) through implicit_svsdf_planner_b200/plan.py — the chain plan_manager.cpp:96-227 drives: map, A* front end,
waypoints and query points, mid-end warm start, SVSDF back end.  Each stage has its own parity tests; this one checks that they fit
together the way the reference wires them."""
import os

import numpy as np
import pytest

from implicit_svsdf_planner_b200 import api, plan, scenes
from oracle import k3_points

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_star_scene_plan_end_to_end():
    g = np.load(os.path.join(HERE, "golden", "map_star_pcd.npz"))
    res, ks = float(g["occupancy_resolution"]), int(g["kernel_size"])
    cmap = plan.gridmap3d_from_cloud(g["points"], res, int(g["sta_threshold"]))
    ctx = api.Context("star", weight_p=scenes.YAML["weight_p"], safety_hor=scenes.YAML["safety_hor"], rho=scenes.YAML["rho"])
    r = plan.generate_traj(ctx, cmap, g["start"][:2], g["end"][:2], kernel_size=ks, traj_parlength=float(g["traj_parlength"]))
    assert r["ok"], r
    path, wps, N = r["path"], r["waypoints"], r["N"]
    # front end: starts / ends in the start / goal cells, moves one cell at a time over free cells
    assert np.all(np.abs(path[0, :2] - g["start"][:2]) <= res) and np.all(np.abs(path[-1, :2] - g["end"][:2]) <= res)
    step = np.abs(np.diff(path[:, :2], axis=0)).max(axis=1)
    assert np.all(step <= res * (1 + 1e-9)) and np.all(step > 0)
    # waypoints and pieces as generateTraj cuts them
    idx, w2 = plan.waypoints_of_path(path, float(g["traj_parlength"]), res)
    assert np.array_equal(w2, wps) and N == len(wps) + 1
    # query points: the device extraction on the 3-D map equals the reference's rule (oracle restatement) for these waypoints
    gm = k3_points.gridmap_from_cloud(g["points"], res, int(g["sta_threshold"]))
    ref_pts = k3_points.query_points(gm, wps, [ks * res / 3.0] * 3)
    assert r["n_points"] == len(ref_pts) > 10
    # mid end -> back end: the back end starts from the warm start and does not end above it
    assert r["mid"]["status"] >= 0 and r["status"] >= 0
    assert np.isfinite(r["final_cost"]) and r["final_cost"] <= r["cost_at_warm_start"] * (1 + 1e-12)
    assert np.all(r["T"] > 0) and r["coeffs"].shape[0] == 18 * N or r["coeffs"].size == 18 * N
    # the optimised spline still starts and ends where the path does
    b = np.asarray(r["coeffs"]).reshape(3, 6 * N).T if np.asarray(r["coeffs"]).ndim == 1 else np.asarray(r["coeffs"])
    assert np.allclose(b[0], path[0], atol=1e-9)
    ctx.close()
