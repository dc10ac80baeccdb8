"""CPU tests of the oracle's restatement of the A* front end's collision kernels (oracle/frontend_oracle.hpp):
BasicShape::initShape, kernelConv (both variants of the reference), visit_kernels_by_distance / checkKernelValue."""
import numpy as np
import pytest

from implicit_svsdf_planner_b200 import batch

SHAPES = ["star", "sdHorseshoe", "sdPie", "sdTunnel", "sdRoundedCross", "sdMoon", "sdUnevenCapsule", "Circle"]


def random_occ(X, Y, density, seed):
    rng = np.random.default_rng(seed)
    return rng.random((X, Y)) < density


def test_circle_kernels_are_the_closed_form(oracle_mod):
    # Circle: radius 1 (Shape.hpp:433-480); cell set iff |centre| - 1 <= max(safeh, res / 2), the same for every yaw
    ks, K, res, safeh = 17, 18, 0.25, 0.0
    yaw, cells, byt = oracle_mod.shape_kernels("Circle", ks, K, res, safeh)
    side = (ks - 1) // 2
    a = (np.arange(ks) - side) * res
    want = np.sqrt(a[:, None] ** 2 + a[None, :] ** 2) - 1.0 <= max(safeh, res / 2)
    assert np.allclose(yaw, -np.pi + 2 * np.pi / K * np.arange(K), atol=1e-14)
    for k in range(K):
        assert np.array_equal(cells[k], want)
    # byte kernels: MSB-first rows (generateByteKernel, Shape.hpp:194-216)
    assert np.array_equal(np.unpackbits(byt, axis=2, bitorder="big")[:, :, :ks].astype(bool), cells)


@pytest.mark.parametrize("shape", SHAPES)
def test_kernels_rotate_with_the_yaw_and_both_conv_variants_agree(oracle_mod, shape):
    ks, K, res = 17, 18, 0.5
    yaw, cells, _ = oracle_mod.shape_kernels(shape, ks, K, res, 0.1)
    assert cells.any() and cells[:, ks // 2, ks // 2].all() or shape in ("sdHorseshoe", "sdMoon", "sdPie", "sdTunnel")
    # kernel k + K/2 is kernel k turned by pi: the cell grid is symmetric under the point reflection (a, b) -> (ks-1-a, ks-1-b)
    assert np.array_equal(cells[K // 2 :], cells[: K // 2, ::-1, ::-1])
    occ = random_occ(50, 61, 0.03, 7)
    a = oracle_mod.cspace(shape, occ, ks, K, res, 0.1, variant="bool")
    b = oracle_mod.cspace(shape, occ, ks, K, res, 0.1, variant="byte")
    assert np.array_equal(a, b)  # kernelConv<false> == kernelConv<true> (sw_manager.hpp:1043-1066 vs 1068-1095)
    # an empty map is free everywhere; a cell under the kernel centre blocks every kernel that covers its own centre
    assert oracle_mod.cspace(shape, np.zeros((20, 20), bool), ks, K, res, 0.1).all()


def test_cspace_is_the_dilation_of_the_map_by_the_kernel(oracle_mod):
    ks, K, res = 9, 8, 1.0
    _, cells, _ = oracle_mod.shape_kernels("star", ks, K, res, 0.0)
    occ = random_occ(40, 37, 0.05, 3)
    free = oracle_mod.cspace("star", occ, ks, K, res, 0.0)
    h = (ks - 1) // 2
    pad = np.zeros((40 + 2 * h, 37 + 2 * h), bool)
    pad[h:-h, h:-h] = occ
    for k in range(K):
        want = np.ones((40, 37), bool)
        for a, b in zip(*np.nonzero(cells[k])):
            want &= ~pad[a : a + 40, b : b + 37]
        assert np.array_equal(free[k], want)


def test_map_kernel_layout_is_the_batch_modes(oracle_mod):
    """The byte variant reads the inflated packed map of generateMapKernel2D — the layout batch.pack_map_kernel produces
    and the GPU library consumes (svsdf_set_map)."""
    occ = random_occ(33, 45, 0.2, 11)
    kern = batch.pack_map_kernel(occ, 17)
    assert kern.shape == (33 + 16, (45 + 16 + 7) // 8)
    assert np.array_equal(batch.unpack_map_kernel(kern, 33, 45, 17), occ)


def test_check_kernel_value_search_order(oracle_mod):
    ks, K, res = 17, 18, 0.5
    occ = random_occ(60, 60, 0.06, 5)
    free = oracle_mod.cspace("star", occ, ks, K, res, 0.0)
    rng = np.random.default_rng(2)
    ind = rng.integers(0, 60, size=(500, 2))
    fy = rng.uniform(-3.1, 3.1, size=500)
    ok, cy = oracle_mod.check_kernel_value("star", occ, fy, ind, ks, K, res, 0.0)
    pi = 3.1415926536
    for q in range(500):
        fi = min(max(int(K * ((fy[q] + pi) / (2 * pi))), 0), K - 1)
        order = [fi]
        for d in range(1, 6):  # breadth-first ring: s, s-1, s+1, s-2, s+2, ... (11 kernels)
            order += [(fi - d) % K, (fi + d) % K]
        hit = [k for k in order if free[k, ind[q, 0], ind[q, 1]]]
        assert ok[q] == bool(hit)
        if hit:
            assert cy[q] == 2 * pi * hit[0] / K - pi


def test_expand_nodes_semantics(oracle_mod):
    # empty map: every in-map neighbour passes, the child keeps the father's yaw bin; out-of-map neighbours fail
    occ = np.zeros((20, 20), bool)
    ij = np.array([[0, 0], [10, 10], [19, 5]])
    fy = np.array([0.3, -2.0, 3.0])
    ok, cy, parts = oracle_mod.expand_nodes("star", occ, ij, fy, origin=(-5.0, 1.0), map_res=1.0)
    assert ok[1].all() and (parts[1] == 7).all()
    assert ok[0].reshape(3, 3)[1:, 1:].all() and not ok[0].reshape(3, 3)[0, :].any() and not ok[0].reshape(3, 3)[:, 0].any()
    pi, K = 3.1415926536, 18
    for q in range(3):
        fi = int(K * ((fy[q] + pi) / (2 * pi)))
        assert np.all(cy[q][ok[q]] == 2 * pi * fi / K - pi)
    # a fully occupied map: nothing passes (cell occupied, kernel hits, obstacle points inside the shape)
    ok, cy, parts = oracle_mod.expand_nodes("star", np.ones((20, 20), bool), ij, fy, map_res=1.0)
    assert not ok.any() and (parts[1] == 0).all()
    # one obstacle right next to the node: the sub-swept-volume test sees it even where the cell itself is free
    occ = np.zeros((30, 30), bool)
    occ[15, 17] = True
    ok, cy, parts = oracle_mod.expand_nodes("star", occ, np.array([[15, 15]]), np.array([0.0]), map_res=1.0)
    assert (parts[0] & 1).all() and not ok[0].all()


def test_astar_paths_are_valid_and_near_optimal(oracle_mod):
    rng = np.random.default_rng(21)
    X, Y = 50, 44
    occ = rng.random((X, Y)) < 0.005
    origin = (-3.0, 7.5)
    st = np.c_[rng.uniform(origin[0] + 1, origin[0] + X - 1, 12), rng.uniform(origin[1] + 1, origin[1] + Y - 1, 12)]
    go = np.c_[rng.uniform(origin[0] + 1, origin[0] + X - 1, 12), rng.uniform(origin[1] + 1, origin[1] + Y - 1, 12)]
    st[0] = [origin[0] - 5.0, origin[1]]  # outside the map: no search
    paths, ex = oracle_mod.astar("star", occ, st, go, origin=origin, map_res=1.0)
    assert paths[0] is None and ex[0] == 0
    found = [p for p in paths if p is not None]
    assert len(found) >= 6
    for q, p in enumerate(paths):
        if p is None:
            continue
        ij = np.floor((p[:, :2] - np.array(origin)) / 1.0).astype(int)
        assert np.array_equal(ij[0], np.floor((st[q] - np.array(origin))).astype(int))
        assert np.array_equal(ij[-1], np.floor((go[q] - np.array(origin))).astype(int))
        step = np.abs(np.diff(ij, axis=0))
        assert step.max() <= 1 and (step.sum(axis=1) >= 1).all()   # 8-connected moves, no repeated cell
        assert not occ[ij[:, 0], ij[:, 1]].any()
        # every edge of the path passes the node test it was generated by
        ok, cy, _ = oracle_mod.expand_nodes("star", occ, ij[:-1], p[:-1, 2], origin=origin, map_res=1.0)
        k = (ij[1:, 0] - ij[:-1, 0] + 1) * 3 + (ij[1:, 1] - ij[:-1, 1] + 1)
        assert ok[np.arange(len(k)), k].all()
        # cost within 0.1 % + of the octile lower bound when the straight route is free, never below it
        d = np.abs(ij[-1] - ij[0])
        lower = np.sqrt(2) * d.min() + (d.max() - d.min())
        cost = np.sqrt((np.diff(ij, axis=0) ** 2).sum(axis=1)).sum()
        assert cost >= lower - 1e-9


def test_product_batch_astar_host_logic_equals_the_oracle(tmp_path):
    """csrc/host/astar.hpp (the lock-step batch search the library runs on top of svsdf_front_expand) and csrc/host/astar_flat.hpp (the same
    bookkeeping on flat arrays and a binary heap, device-compilable) compiled for the host with the oracle's node test plugged in: paths
    and expansion counts must equal the oracle's literal AstarPathSearch."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "astar_host")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-mfma", "-ffp-contract=off",
                           os.path.join(root, "tests", "cpp", "astar_host_main.cpp"), "-o", exe])
    out = subprocess.run([exe, "40"], capture_output=True, text=True)
    lines = out.stdout.strip().split("\n")
    assert out.returncode == 0 and lines[-1].startswith("OK"), out.stdout + out.stderr
    # the container-free bookkeeping a device-side search would run (csrc/host/astar_flat.hpp) gives the same paths and counts
    assert any(l.startswith("flat bookkeeping: identical") for l in lines), out.stdout
