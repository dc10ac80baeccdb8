"""oracle/k3_points.py — TEST INFRASTRUCTURE ONLY (never imported by the product; tests/ compare the device kernels with it).

numpy / pure-Python restatement of how the reference turns a point cloud into the optimiser's query points:

  PCSmapManager::rcvGlobalMapHandler      src/map_manager/src/PCSmap_manager.cpp:113-190   cloud -> boundary -> occupancy grid
  GridMap3D::createGridMap                src/map_manager/src/Gridmap3D.cpp:25-41          sizes = ceil(extent / resolution)
  GridMap3D::isInMap / getGridIndex       Gridmap3D.cpp:43-75, 137-174                     (0,0,0) outside; floor + clamp inside
  GridMap3D::getGridCubeCenter            Gridmap3D.cpp:181-193                            (index + 0.5) * res + boundary_min
  PCSmapManager::projInMap / unifiedID    map_manager/include/map_manager/PCSmap_manager.h:118-135
  PCSmapManager::getPointsInAABBOutOfLastOne                       PCSmap_manager.h:184-219
  PCSmapManager::generateMapKernel / generateMapKernel2D           PCSmap_manager.h:39-108   byte-packed maps
  PlannerManager::generateTraj (waypoint / box loop)               src/plan_manager/src/plan_manager.cpp:131-175

Everything is three-dimensional as in the reference: a voxel of every occupied z layer inside a waypoint's box becomes a query point,
so a column with several occupied layers yields several points with the same (x, y) — the cost loop zeroes z
(back_end_optimizer.hpp:791), i.e. such a column simply counts several times.  The first waypoint's "last box" is the box around
tmp_pos = (999, 999, 999) (plan_manager.cpp:152): projected into the map it is the single voxel in the map's far corner, which is
therefore skipped for the first waypoint.  aabb_points is an unordered_map keyed by unifiedID: duplicates collapse, the iteration
order is unspecified — results here are sorted by unifiedID, tests compare as sets / in the product's documented order.

Pinned by the reference's own scene: pcds/map_star.pcd (210 points) at config/star.yaml's occupancy_resolution 1.0 and
sta_threshold 1 gives the 148 occupied voxels SURVEY.md §6 quotes (tests/golden/map_star_pcd.npz, tests/test_oracle_k3.py).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass
class GridMap3D:
    boundary_min: np.ndarray  # boundary_xyzmin
    boundary_max: np.ndarray  # boundary_xyzmax
    res: float
    occ: np.ndarray           # [X, Y, Z] bool

    @property
    def size(self):
        return self.occ.shape

    def is_in_map(self, p) -> bool:  # Gridmap3D.cpp:43-75
        return bool(np.all(p >= self.boundary_min) and np.all(p <= self.boundary_max))

    def grid_index(self, p):  # Gridmap3D.cpp:137-174
        p = np.asarray(p, dtype=np.float64)
        if not self.is_in_map(p):
            return (0, 0, 0)
        idx = [int(math.floor((p[a] - self.boundary_min[a]) / self.res)) for a in range(3)]
        X, Y, Z = self.size
        # the clamps as written: `if (iy < 0) ix = 0;` and `if (iz < 0) ix = 0;` reset ix, not iy / iz — unreachable after isInMap
        if idx[0] < 0: idx[0] = 0
        if idx[0] >= X: idx[0] = X - 1
        if idx[1] < 0: idx[0] = 0
        if idx[1] >= Y: idx[1] = Y - 1
        if idx[2] < 0: idx[0] = 0
        if idx[2] >= Z: idx[2] = Z - 1
        return tuple(idx)

    def cube_center(self, i, j, k):  # Gridmap3D.cpp:181-193
        X, Y, Z = self.size
        if not (0 <= i < X and 0 <= j < Y and 0 <= k < Z):
            return np.zeros(3)
        return np.array([(i + 0.5) * self.res, (j + 0.5) * self.res, (k + 0.5) * self.res]) + self.boundary_min

    def index_of_center(self, c):
        """Inverse of cube_center (NOT getGridIndex: the last cell of an axis sticks out of boundary_max, where getGridIndex
        answers (0, 0, 0))."""
        return tuple(int(round((c[a] - self.boundary_min[a]) / self.res - 0.5)) for a in range(3))

    def unified_id(self, i, j, k) -> int:  # PCSmap_manager.h:118-125
        X, Y, _ = self.size
        return k * X * Y + j * X + i

    def proj_in_map(self, p):  # PCSmap_manager.h:128-135
        return np.minimum(np.maximum(np.asarray(p, dtype=np.float64), self.boundary_min), self.boundary_max)


def gridmap_from_cloud(points_f32: np.ndarray, res: float, sta_threshold: int = 1) -> GridMap3D:
    """rcvGlobalMapHandler: boundary = min / max of the cloud (pcl::PointXYZ floats widened to double), counts per voxel,
    occupied where count >= sta_threshold."""
    pts = np.asarray(points_f32, dtype=np.float32).astype(np.float64).reshape(-1, 3)
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    size = [int(math.ceil((hi[a] - lo[a]) / res)) for a in range(3)]  # createGridMap
    gm = GridMap3D(boundary_min=lo, boundary_max=hi, res=float(res), occ=np.zeros(tuple(max(s, 0) for s in size), dtype=bool))
    cnt = np.zeros(gm.size, dtype=np.int64)
    if cnt.size:
        for p in pts:
            cnt[gm.grid_index(p)] += 1
    gm.occ = cnt >= sta_threshold
    return gm


def points_in_aabb_out_of_last_one(gm: GridMap3D, center, center_last, half, aabb: dict) -> None:
    """PCSmap_manager.h:184-219: occupied voxels of the box around `center` that are outside the box around `center_last`, into the
    map id -> centre (emplace: the first insertion wins; the value is the same anyway)."""
    half = np.asarray(half, dtype=np.float64) * np.ones(3)
    c1 = gm.grid_index(gm.proj_in_map(np.asarray(center, dtype=np.float64) - half))
    c2 = gm.grid_index(gm.proj_in_map(np.asarray(center, dtype=np.float64) + half))
    l1 = gm.grid_index(gm.proj_in_map(np.asarray(center_last, dtype=np.float64) - half))
    l2 = gm.grid_index(gm.proj_in_map(np.asarray(center_last, dtype=np.float64) + half))
    X, Y, Z = gm.size
    for i in range(c1[0], c2[0] + 1):
        for j in range(c1[1], c2[1] + 1):
            for k in range(c1[2], c2[2] + 1):
                if i > l2[0] or i < l1[0] or j > l2[1] or j < l1[1] or k > l2[2] or k < l1[2]:
                    if 0 <= i < X and 0 <= j < Y and 0 <= k < Z and gm.occ[i, j, k]:  # isIndexOccupied: invalid index -> false
                        aabb.setdefault(gm.unified_id(i, j, k), gm.cube_center(i, j, k))


TMP_POS = np.array([999.0, 999.0, 999.0])  # plan_manager.cpp:152


def query_points(gm: GridMap3D, waypoints, half) -> np.ndarray:
    """plan_manager.cpp:156-175: one getPointsInAABBOutOfLastOne per waypoint, the previous waypoint as centre of the last box
    (tmp_pos for the first).  Returns the voxel centres [n, 3] sorted by unifiedID."""
    aabb: dict = {}
    last = TMP_POS
    for wp in np.asarray(waypoints, dtype=np.float64).reshape(-1, 3):
        points_in_aabb_out_of_last_one(gm, wp, last, half, aabb)
        last = wp
    if not aabb:
        return np.zeros((0, 3))
    return np.stack([aabb[k] for k in sorted(aabb)])


def waypoints_of_path(path, traj_parlength: float, res: float):
    """plan_manager.cpp:131-158: every index_gap-th cell of the front-end path (ends excluded) becomes a waypoint."""
    path = np.asarray(path, dtype=np.float64).reshape(-1, 3)
    n = path.shape[0]
    t = traj_parlength
    gap = int(math.ceil(t / res))
    while gap >= n - 1:
        t /= 1.5
        gap = int(math.ceil(t / res))
    return path[gap:n - 1:gap]


# ---- byte-packed maps (what the reference hands its front end; the product's device kernels read these layouts) ----------------
OR_MASK = np.array([0x80, 0x40, 0x20, 0x10, 0x08, 0x04, 0x02, 0x01], dtype=np.uint8)  # Shape.hpp / PCSmap_manager.h or_mask


def generate_map_kernel(gm: GridMap3D, kernel_size: int) -> np.ndarray:
    """generateMapKernel (PCSmap_manager.h:39-78): [(X + 2h)][(Y + 2h)][ceil((Z + 2h) / 8)] bytes, z bits MSB first."""
    h = (kernel_size - 1) // 2
    X, Y, Z = gm.size
    bl = (Z + 2 * h + 7) // 8
    out = np.zeros((X + 2 * h, Y + 2 * h, bl), dtype=np.uint8)
    for x, y, z in zip(*np.nonzero(gm.occ)):
        fz = z + h
        out[x + h, y + h, fz // 8] |= OR_MASK[fz % 8]
    return out


def generate_map_kernel_2d(gm: GridMap3D, kernel_size: int) -> np.ndarray:
    """generateMapKernel2D (PCSmap_manager.h:81-108): the z = 0 layer, [(X + 2h)][ceil((Y + 2h) / 8)] bytes, y bits MSB first."""
    h = (kernel_size - 1) // 2
    X, Y, _ = gm.size
    bl = (Y + 2 * h + 7) // 8
    out = np.zeros((X + 2 * h, bl), dtype=np.uint8)
    for x, y in zip(*np.nonzero(gm.occ[:, :, 0])):
        fy = y + h
        out[x + h, fy // 8] |= OR_MASK[fy % 8]
    return out


# ---- the flat (z = 0 layer) case the batch mode uses ------------------------------------------------------------------------------
def gridmap_2d(occ2d: np.ndarray, origin_xy, res: float) -> GridMap3D:
    """A one-layer map: X x Y x 1 voxels, boundary z = [0, res]."""
    occ2d = np.asarray(occ2d, dtype=bool)
    X, Y = occ2d.shape
    lo = np.array([origin_xy[0], origin_xy[1], 0.0])
    hi = lo + np.array([X, Y, 1]) * res
    return GridMap3D(boundary_min=lo, boundary_max=hi, res=float(res), occ=occ2d[:, :, None].copy())


def query_points_2d(occ2d, origin_xy, res, waypoints_xy, half, keepout=None, clearance: float = 0.0) -> np.ndarray:
    """query_points on a one-layer map, as [n, 3] with z = 0, in ascending (i * Y + j) order (the order the device kernel emits).
    keepout / clearance: the synthetic scenes' option (not in the reference): drop cells within `clearance` of a keep-out sample."""
    gm = gridmap_2d(occ2d, origin_xy, res)
    wp = np.zeros((len(waypoints_xy), 3))
    wp[:, :2] = np.asarray(waypoints_xy, dtype=np.float64)[:, :2]
    wp[:, 2] = 0.5 * res
    aabb: dict = {}
    last = TMP_POS
    for w in wp:
        points_in_aabb_out_of_last_one(gm, w, last, [half, half, half], aabb)
        last = w
    X, Y, _ = gm.size
    ids = sorted(aabb)
    ij = sorted(((k % X), (k // X) % Y) for k in ids)  # (i, j) ascending i * Y + j
    pts = np.zeros((len(ij), 3))
    for n, (i, j) in enumerate(ij):
        pts[n, :2] = gm.cube_center(i, j, 0)[:2]
    if keepout is not None and len(keepout) and len(pts):
        ko = np.asarray(keepout, dtype=np.float64).reshape(-1, 2)
        dx = pts[:, None, 0] - ko[None, :, 0]
        dy = pts[:, None, 1] - ko[None, :, 1]
        pts = pts[(dx * dx + dy * dy > clearance * clearance).all(axis=1)]
    return pts
