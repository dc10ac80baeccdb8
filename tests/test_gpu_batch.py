"""GPU tests of the batch C ABI (svsdf_optimize_batch / svsdf_cost_grad_batch): a pool of contexts on one GPU solves
independent problems concurrently; results equal the one-at-a-time calls bit for bit whatever context solved what."""
import numpy as np
import pytest

from implicit_svsdf_planner_b200 import api, batch, scenes

pytestmark = pytest.mark.gpu


def _problems(n, P=3000):
    out = []
    for k in range(n):
        sc = scenes.make_scene("star", 8, P, seed_traj=scenes.SEED_TRAJ + 5 * k, seed_map=scenes.SEED_MAP + 5 * k)
        out.append(sc)
    return out


def test_cost_grad_batch_equals_single_calls():
    scs = _problems(5)
    ctxs = [api.Context("star") for _ in range(3)]
    T = np.stack([sc.T for sc in scs])
    co = np.stack([sc.coeffs_colmajor() for sc in scs])
    rc, cost, gT, gC = api.cost_grad_batch(ctxs, [sc.points for sc in scs], T, co, 8)
    assert rc == 0
    one = api.Context("star")
    for k, sc in enumerate(scs):
        one.set_points(sc.points)
        c, t, g = one.cost_grad(sc.T, sc.coeffs_colmajor())
        assert c == cost[k] and np.array_equal(t, gT[k]) and np.array_equal(g, gC[k])


def test_optimize_batch_equals_sequential_optimize_and_uses_the_queue():
    scs = _problems(6, P=2000)
    params = api.default_lbfgs_params(mem_size=16, past=3, delta=1e-6, g_epsilon=0.0, max_iterations=6, min_step=1e-32)
    descr = [dict(init_s=sc.init_s, final_s=sc.final_s, x0=sc.x0, points=sc.points) for sc in scs]
    ctxs = [api.Context("star") for _ in range(3)]
    q = batch.WorkQueue(len(scs), order=np.arange(len(scs))[::-1])
    rc, X, st, status, npts = api.optimize_batch(ctxs, descr, 8, params, next_index=q.next)
    assert rc == 0 and sorted(q.taken) == list(range(len(scs))) and list(npts) == [sc.P for sc in scs]
    one = api.Context("star")
    for k, sc in enumerate(scs):
        one.set_points(sc.points)
        rc1, x1, T1, b1, st1 = one.optimize(sc.init_s, sc.final_s, sc.x0, 8, params)
        assert rc1 == status[k]
        assert np.array_equal(x1, X[k]) and st1["final_cost"] == st[k]["final_cost"] and st1["evaluations"] == st[k]["evaluations"]
    # without a queue callback the pool hands the problems out itself
    rc, X2, st2, status2, _ = api.optimize_batch(ctxs[:2], descr, 8, params)
    assert rc == 0 and np.array_equal(X2, X) and np.array_equal(status2, status)


def test_optimize_batch_with_device_side_point_extraction():
    gm = batch.make_random_map(extent=40.0, res=0.1, density=0.3, seed=5)
    X = Y = gm.occ.shape[0]
    ks = 17
    kern = batch.pack_map_kernel(gm.occ, ks)
    ctxs = [api.Context("star") for _ in range(2)]
    for c in ctxs:
        c.set_map(kern, X, Y, ks, gm.origin, gm.res)
    probs = scenes.make_batch_problems(4, seed=3, extent=(6.0, 34.0), coords_path=None)
    half = 17 * 1.0 / 3.0
    descr = []
    for k, sg in enumerate(probs):
        init_s, final_s, q, T = scenes.make_trajectory("star", 8, 100 + k, sg[:2], sg[2:4])
        b = scenes.minco_dense(init_s, final_s, q, T)
        wps = np.concatenate([init_s[:2, :1], q[:2], final_s[:2, :1]], axis=1).T
        descr.append(dict(init_s=init_s, final_s=final_s, x0=np.concatenate([scenes.backward_T(T), q.T.reshape(-1)]), waypoints=wps,
                          half=half, keepout=batch.keepout_samples(b, T), clearance=2.75))
    params = api.default_lbfgs_params(mem_size=16, past=3, delta=1e-6, g_epsilon=0.0, max_iterations=4, min_step=1e-32)
    rc, Xo, st, status, npts = api.optimize_batch(ctxs, descr, 8, params)
    assert rc == 0 and (npts > 100).all()
    # the same through the single-problem calls
    one = ctxs[0]
    for k, d in enumerate(descr):
        P = one.extract_points(d["waypoints"], d["half"], d["keepout"], d["clearance"])
        assert P == npts[k]
        rc1, x1, _, _, st1 = one.optimize(d["init_s"], d["final_s"], d["x0"], 8, params)
        assert np.array_equal(x1, Xo[k]) and rc1 == status[k]
