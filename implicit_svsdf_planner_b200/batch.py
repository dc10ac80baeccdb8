"""Batch-of-problems mode (BASELINE config 5): independent start/goal problems sharded over the GPUs of one box.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  Problems are independent — own
decision vector, own query set, own optimiser state — so the shard is a contiguous block of problem indices per rank
and there is NO collective on the data path.  The only shared datum is the map, broadcast once from rank 0 as the
reference's bit-packed 2-D "map kernel" (src/map_manager/include/map_manager/PCSmap_manager.h:32, 81-108:
(X + 2h) x ceil((Y + 2h) / 8) bytes, h = (kernel_size - 1) / 2, MSB-first); results are gathered at the end.

Query points of a problem are the occupied cell centres inside the AABBs (half-size kernel_size * res / 3) around its
waypoints, visiting for each waypoint only the cells outside the previous waypoint's box and de-duplicating by cell
id — the 2-D (z = 0 layer) restatement of plan_manager.cpp:156-175 / PCSmap_manager.h:118-125, 184-219.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional

import os

import numpy as np

from . import scenes

OR_MASK = np.array([0x80, 0x40, 0x20, 0x10, 0x08, 0x04, 0x02, 0x01], dtype=np.uint8)  # PCSmap_manager.h:32


@dataclasses.dataclass
class GridMap:
    occ: np.ndarray  # (X, Y) bool, z = 0 layer
    origin: np.ndarray  # (2,) world coordinates of the corner of cell (0, 0)  (boundary_xyzmin)
    res: float

    @property
    def shape(self):
        return self.occ.shape

    def cell_centers(self, ij: np.ndarray) -> np.ndarray:
        """GridMap3D::getGridCubeCenter: (idx + 0.5) * res + min."""
        return (ij + 0.5) * self.res + self.origin[None, :]


def pack_map_kernel(occ: np.ndarray, kernel_size: int = 17) -> np.ndarray:
    """generateMapKernel2D (PCSmap_manager.h:81-108)."""
    X, Y = occ.shape
    h = (kernel_size - 1) // 2
    xs, ys = X + 2 * h, Y + 2 * h
    row = (ys + 7) // 8
    out = np.zeros((xs, row), dtype=np.uint8)
    ii, jj = np.nonzero(occ)
    fx, fy = ii + h, jj + h
    np.bitwise_or.at(out, (fx, fy // 8), OR_MASK[fy % 8])
    return out


def unpack_map_kernel(kernel: np.ndarray, X: int, Y: int, kernel_size: int = 17) -> np.ndarray:
    h = (kernel_size - 1) // 2
    bits = np.unpackbits(kernel, axis=1, bitorder="big")  # MSB first
    return bits[h : h + X, h : h + Y].astype(bool)


def pack_map_kernel_from_points(points: np.ndarray, res: float, kernel_size: int = 17) -> np.ndarray:
    """Occupancy grid from a point set (bounding box of the points, PCSmap_manager.cpp:116-150), then packed."""
    gm = gridmap_from_points(points, res)
    return pack_map_kernel(gm.occ, kernel_size)


def gridmap_from_points(points: np.ndarray, res: float) -> GridMap:
    xy = np.asarray(points, dtype=np.float64)[:, :2]
    lo = xy.min(axis=0) - 0.5 * res
    hi = xy.max(axis=0) + 0.5 * res
    size = np.maximum(np.ceil((hi - lo) / res).astype(int), 1)  # Gridmap3D.cpp:25-41
    ij = np.clip(np.floor((xy - lo) / res).astype(int), 0, size - 1)
    occ = np.zeros(tuple(size), dtype=bool)
    occ[ij[:, 0], ij[:, 1]] = True
    return GridMap(occ=occ, origin=lo, res=res)


def make_random_map(extent: float = 60.0, res: float = 0.025, density: float = 0.3, seed: int = scenes.SEED_MAP) -> GridMap:
    rng = np.random.Generator(np.random.MT19937(seed))
    n = int(np.ceil(extent / res))
    occ = rng.random((n, n), dtype=np.float32) < density
    return GridMap(occ=occ, origin=np.zeros(2), res=res)


def bind_to_gpu_numa(device_index: int) -> Optional[int]:
    """Pin this process (and the threads it starts later) to the CPUs the driver reports as local to GPU `device_index`
    (NVML's ideal CPU affinity: the NUMA node the GPU hangs off), so that the pinned staging buffers are first-touched in that
    node's memory and the H2D copies do not cross sockets.  For multi-rank launches (one process per GPU); returns the number
    of CPUs in the mask, or None when NVML / sched_setaffinity is unavailable (nothing is changed then)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def partition(n_problems: int, world: int, rank: int) -> range:
    """Contiguous block partition (remainder spread over the first ranks)."""
    base, rem = divmod(n_problems, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def broadcast_map(kernel: Optional[np.ndarray], device=None, src: int = 0):
    """Broadcast the packed map kernel from `src` to every rank (the only collective of the batch mode).
    Returns the kernel as a uint8 torch tensor on `device` (CPU for gloo)."""
    import torch
    import torch.distributed as dist

    device = device if device is not None else torch.device("cpu")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return torch.from_numpy(np.ascontiguousarray(kernel)).to(device)
    shape = torch.zeros(2, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        shape = torch.tensor(list(kernel.shape), dtype=torch.int64, device=device)
    dist.broadcast(shape, src=src)
    if dist.get_rank() == src:
        buf = torch.from_numpy(np.ascontiguousarray(kernel)).to(device)
    else:
        buf = torch.empty(tuple(int(v) for v in shape.tolist()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src=src)
    return buf


def problem_scene(gm: GridMap, start, goal, extract: Callable, N: int = 8, P: Optional[int] = None, seed: int = 0,
                  clearance: float = 2.75) -> scenes.Scene:
    """One start/goal problem on the shared map: seeded nominal spline, query points taken from the map around its waypoints by
    `extract(gm, waypoints_xy, half, keepout_xy, clearance) -> [n, 3]` (on a GPU box: svsdf_extract_points through
    api.Context, see scripts/run_batch.py; the CPU tests pass the oracle's restatement), a corridor around the nominal path
    kept free (stand-in for the A* feasibility of the reference's front end), optionally sub-sampled to exactly P points."""
    init_s, final_s, q, T = scenes.make_trajectory("star", N, seed, start, goal)
    b = scenes.minco_dense(init_s, final_s, q, T)
    half = scenes.YAML["kernel_size"] * scenes.YAML["occupancy_resolution"] / 3.0
    wps = np.concatenate([init_s[:2, :1], q[:2], final_s[:2, :1]], axis=1).T
    pts = extract(gm, wps, half, keepout_samples(b, T), clearance)
    if P is not None and pts.shape[0] > P:
        rng = np.random.Generator(np.random.MT19937(seed + 1))
        pts = pts[np.sort(rng.choice(pts.shape[0], size=P, replace=False))]
    return scenes.Scene(shape="star", N=N, init_s=init_s, final_s=final_s, q=q, T=T, coeffs=b, points=pts, resolution=gm.res)


def keepout_samples(b: np.ndarray, T: np.ndarray, n: int = 150) -> np.ndarray:
    """Samples of the nominal path used as keep-out centres (synthetic stand-in for the A* front end's guarantee that the
    initial path is collision free; <= 160 samples, the device kernel's limit)."""
    return scenes.eval_traj_xy(b, T, np.linspace(0.0, float(T.sum()), n))[:, :2]


def lpt_order(problems: np.ndarray) -> np.ndarray:
    """Longest-processing-time-first order for the dynamic queue: longer start-goal distance = more query points and more
    optimiser work, so long problems are handed out first and the short ones fill the tail."""
    d = np.asarray(problems)[:, 2:4] - np.asarray(problems)[:, 0:2]
    return np.argsort(-np.linalg.norm(d, axis=1), kind="stable")


class WorkQueue:
    """Problem indices handed out one at a time to whoever asks next — across every rank of the job (iteration
    counts differ per problem, so a static split leaves GPUs idle).  The counter lives in the torch.distributed store
    (an atomic add on the rank-0 store server, one TCP round trip per problem — microseconds against the ~0.1 s of a
    problem); without a process group it is a local counter.  `order` maps the k-th ticket to a problem index (LPT)."""

    def __init__(self, n: int, order=None, key: str = "svsdf_queue"):
        import threading

        self.n = int(n)
        self.order = np.arange(self.n) if order is None else np.asarray(order)
        self.key = key
        self.taken = []
        self._lock = threading.Lock()
        self._local = 0
        self.store = None
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self.store = dist.distributed_c10d._get_default_store()
        except Exception:
            self.store = None

    def next(self) -> int:
        """Next problem index, or -1 when the queue is empty.  Thread-safe (called from the pool's worker threads)."""
        with self._lock:
            if self.store is not None:
                t = int(self.store.add(self.key, 1)) - 1
            else:
                t = self._local
                self._local += 1
            if t >= self.n:
                return -1
            k = int(self.order[t])
            self.taken.append(k)
            return k


def gather_rows(local: np.ndarray, mine, device=None) -> np.ndarray:
    """Every rank contributes the rows it solved (`mine`: their indices); returns the full table on every rank.  Ownership
    comes from the indices, not from the values (a solver result may legitimately contain NaN)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    dev = device if device is not None else torch.device("cpu")
    mask = np.zeros(local.shape[0])
    mask[list(mine)] = 1.0
    vals = np.where(mask[:, None] > 0, local, 0.0)
    nan_mask = np.isnan(vals)
    t = torch.from_numpy(np.where(nan_mask, 0.0, vals)).to(dev)
    nn = torch.from_numpy(nan_mask.astype(np.float64)).to(dev)
    m = torch.from_numpy(mask).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(nn, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.SUM)
    cnt = m.cpu().numpy()
    if not np.all(cnt == 1.0):
        raise RuntimeError(f"every problem must be solved by exactly one rank (counts: min {cnt.min()}, max {cnt.max()})")
    out = t.cpu().numpy()
    out[nn.cpu().numpy() > 0] = np.nan
    return out


class BatchRunner:
    """Runs the problems of this rank and gathers the per-problem results.

    solve(scene, index) -> 1-D float array (fixed length) is injected: on a GPU box it wraps
    ``api.Context.optimize``; the gloo tests inject a CPU stand-in so the sharding / broadcast / gather logic is
    exercised without CUDA.  dynamic=False: contiguous static split (`partition`); dynamic=True: `WorkQueue` shared by all
    ranks, LPT order."""

    def __init__(self, solve: Callable[[scenes.Scene, int], np.ndarray], result_len: int, extract: Callable, dynamic: bool = False):
        self.solve = solve
        self.result_len = result_len
        self.extract = extract  # query-point construction, see problem_scene
        self.dynamic = dynamic
        self.mine = []

    def run(self, gm: GridMap, problems: np.ndarray, N: int = 8, P: Optional[int] = None, device=None):
        import torch.distributed as dist

        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        n = problems.shape[0]
        local = np.zeros((n, self.result_len))

        def one(k):
            sg = problems[k]
            sc = problem_scene(gm, sg[:2], sg[2:4], self.extract, N=N, P=P, seed=scenes.SEED_BATCH + k)
            local[k] = self.solve(sc, k)

        if self.dynamic:
            q = WorkQueue(n, lpt_order(problems))
            if world > 1:
                dist.barrier()  # nobody draws before every rank has built its queue object
            while True:
                k = q.next()
                if k < 0:
                    break
                one(k)
            self.mine = list(q.taken)
        else:
            self.mine = list(partition(n, world, rank))
            for k in self.mine:
                one(k)
        return gather_rows(local, self.mine, device)
