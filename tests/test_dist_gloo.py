"""CPU, world_size = 2 over gloo: the host-side logic of the batch-of-problems mode (map packing + broadcast,
contiguous sharding with no data-path collective, result gather).  The per-problem solver is a CPU stand-in injected
into BatchRunner (the real one wraps the CUDA context and is exercised by the GPU bench)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from implicit_svsdf_planner_b200 import batch, scenes  # noqa: E402
from oracle import k3_points  # noqa: E402  (the CPU restatement of the reference's point collection: the checker, tests only)


def _extract(gm, wps, half, keepout=None, clearance=0.0):
    return k3_points.query_points_2d(gm.occ, gm.origin, gm.res, wps, half, keepout, clearance)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_solve(scene, k):
    # deterministic function of the problem's inputs only (what a real solver's result would depend on)
    return np.array([k, scene.P, float(scene.points[:, :2].sum()), float(scene.q.sum())])


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gm = batch.make_random_map(extent=60.0, res=0.5, density=0.3, seed=123) if rank == 0 else None
        kern = batch.pack_map_kernel(gm.occ, 17) if rank == 0 else None
        t = batch.broadcast_map(kern)  # the only collective before the run
        X = Y = 120
        occ = batch.unpack_map_kernel(t.numpy(), X, Y, 17)
        gm_local = batch.GridMap(occ=occ, origin=np.zeros(2), res=0.5)
        problems = scenes.make_batch_problems(7, seed=99)
        runner = batch.BatchRunner(_fake_solve, 4, _extract)
        out = runner.run(gm_local, problems, N=8, P=None)
        np.save(os.path.join(tmpdir, f"out_{rank}.npy"), out)
        np.save(os.path.join(tmpdir, f"occ_{rank}.npy"), occ)
        mine = list(batch.partition(7, world, rank))
        np.save(os.path.join(tmpdir, f"mine_{rank}.npy"), np.array(mine))
        # dynamic work queue shared by the ranks (counter in the c10d store), LPT order
        dyn = batch.BatchRunner(_fake_solve, 4, _extract, dynamic=True)
        out_d = dyn.run(gm_local, problems, N=8, P=None)
        np.save(os.path.join(tmpdir, f"outdyn_{rank}.npy"), out_d)
        np.save(os.path.join(tmpdir, f"minedyn_{rank}.npy"), np.array(dyn.mine, dtype=np.int64))
    finally:
        dist.destroy_process_group()


def test_partition_is_a_contiguous_cover():
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 3, 8):
            parts = [list(batch.partition(n, w, r)) for r in range(w)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_map_kernel_packing_matches_reference_layout():
    rng = np.random.default_rng(0)
    occ = rng.random((13, 21)) < 0.4
    k = batch.pack_map_kernel(occ, 17)
    h = 8
    assert k.shape == (13 + 2 * h, (21 + 2 * h + 7) // 8) and k.dtype == np.uint8
    # bit (x, y) lives in byte [(x+h), (y+h)//8] under mask 0x80 >> ((y+h) % 8)   (PCSmap_manager.h:32, 96-103)
    for x, y in ((0, 0), (5, 7), (12, 20), (3, 15)):
        bit = (k[x + h, (y + h) // 8] >> (7 - (y + h) % 8)) & 1
        assert bool(bit) == bool(occ[x, y])
    assert np.array_equal(batch.unpack_map_kernel(k, 13, 21, 17), occ)
    assert int(np.unpackbits(k).sum()) == int(occ.sum())  # the padding ring stays empty


def test_query_point_extraction_dedups_and_skips_previous_box():
    occ = np.ones((40, 40), dtype=bool)
    gm = batch.GridMap(occ=occ, origin=np.zeros(2), res=1.0)
    half = 3.0
    one = _extract(gm, np.array([[10.5, 10.5]]), half)
    assert one.shape[0] == 7 * 7 and np.all(one[:, 2] == 0)
    two = _extract(gm, np.array([[10.5, 10.5], [12.5, 10.5]]), half)
    assert two.shape[0] == 7 * 7 + 2 * 7  # only the two new columns of the second box
    assert np.unique(two[:, :2], axis=0).shape[0] == two.shape[0]
    # cell centres: min + (idx + 0.5) * res
    assert np.allclose(np.modf(one[:, :2])[0], 0.5)


def test_world_size_2_gloo_broadcast_shard_gather(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    o0, o1 = np.load(tmp_path / "out_0.npy"), np.load(tmp_path / "out_1.npy")
    assert np.array_equal(o0, o1) and not np.isnan(o0).any()  # every rank ends with all results
    assert np.array_equal(np.load(tmp_path / "occ_0.npy"), np.load(tmp_path / "occ_1.npy"))  # map arrived intact
    m0, m1 = np.load(tmp_path / "mine_0.npy"), np.load(tmp_path / "mine_1.npy")
    assert list(m0) == [0, 1, 2, 3] and list(m1) == [4, 5, 6]
    # multi-rank result == single-process result (problems are independent)
    gm = batch.make_random_map(extent=60.0, res=0.5, density=0.3, seed=123)
    problems = scenes.make_batch_problems(7, seed=99)
    ref = batch.BatchRunner(_fake_solve, 4, _extract).run(gm, problems, N=8, P=None)
    assert np.array_equal(ref, o0)
    assert np.array_equal(o0[:, 0], np.arange(7))
    # dynamic queue: same table on both ranks, every problem handed out exactly once across the ranks, first ticket = longest
    d0, d1 = np.load(tmp_path / "outdyn_0.npy"), np.load(tmp_path / "outdyn_1.npy")
    assert np.array_equal(d0, d1) and np.array_equal(d0, ref)
    t0, t1 = list(np.load(tmp_path / "minedyn_0.npy")), list(np.load(tmp_path / "minedyn_1.npy"))
    assert sorted(t0 + t1) == list(range(7))
    order = list(batch.lpt_order(problems))
    assert order[0] in (t0[:1] + t1[:1])


def test_config5_problem_set_uses_the_reference_coords_file():
    sg = scenes.make_batch_problems(1200)
    c = np.loadtxt(scenes.COORDS_TXT, delimiter=",")
    assert c.shape == (1000, 6)  # src/coords.txt of the reference (copied as a fixture)
    L = np.linalg.norm(sg[:, 2:] - sg[:, :2], axis=1)
    assert (L >= 25.0 - 1e-9).mean() > 0.97 and sg.min() >= 2.0 and sg.max() <= 58.0
    keep = np.linalg.norm(c[:, 3:5] - c[:, 0:2], axis=1) >= 25.0
    inside = ((c[:, [0, 1, 3, 4]] >= 2.0) & (c[:, [0, 1, 3, 4]] <= 58.0)).all(axis=1)
    ok = keep & inside
    assert ok.sum() > 300 and np.array_equal(sg[:1000][ok], c[ok][:, [0, 1, 3, 4]])  # untouched rows are the file's rows
    assert np.array_equal(sg, scenes.make_batch_problems(1200))  # deterministic
