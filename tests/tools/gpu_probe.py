"""Ad-hoc GPU probe (development aid): parity vs oracle on small scenes + raw kernel timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from implicit_svsdf_planner_b200 import api, scenes
from oracle import oracle_py as O

def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(1e-300, np.linalg.norm(b))

for strict in (False, True):
    for (P, cl) in ((2000, 2.75), (400, 2.35)):
        sc = scenes.make_scene("star", 8, P, clearance=cl)
        ctx = api.Context("star", strict_fp=strict)
        ctx.set_points(sc.points)
        co = sc.coeffs_colmajor()
        pts = np.c_[sc.points[:, :2], np.zeros(sc.P)]
        orc = O.Oracle("star", threads=O.num_procs()); orc.set_points(sc.points); orc.set_traj(sc.T, co)
        so, to, go = orc.query_outer(pts)
        sg, tg, gg, _ = ctx.query(sc.T, co, pts, outer_only=True)
        print(f"strict={strict} P={P}: outer sdf maxabs {np.abs(sg-so).max():.3e} t* maxabs {np.abs(tg-to).max():.3e} grad maxabs {np.abs(gg-go).max():.3e}")
        so, to, go, ro = orc.query(pts)
        sg, tg, gg, rg = ctx.query(sc.T, co, pts, outer_only=False)
        ins = so <= 0
        print(f"   true: inside {ins.sum()} sdf maxabs {np.abs(sg-so).max():.3e} t* maxabs {np.abs(tg-to).max():.3e} grad maxabs {np.abs(gg-go).max():.3e} rounds eq {np.array_equal(ro, rg)}")
        c0, gT0, gC0, pp, inside = orc.cost_grad(sc.T, co)
        c1, gT1, gC1 = ctx.cost_grad(sc.T, co)
        print(f"   cost cpu {c0:.12g} gpu {c1:.12g} rel {abs(c1-c0)/abs(c0):.3e} gradC rel {rel(gC1,gC0):.3e} gradT rel {rel(gT1,gT0):.3e}")
        c2, gT2, gC2 = ctx.cost_grad(sc.T, co)
        print(f"   deterministic: {c1==c2 and np.array_equal(gC1,gC2) and np.array_equal(gT1,gT2)}")

ctx = api.Context("star")
print("fp64 peak TFLOP/s:", ctx.fp64_peak_tflops())
for P in (2000, 200000):
    sc = scenes.make_scene("star", 8, P)
    ctx.set_points(sc.points)
    co = sc.coeffs_colmajor()
    ms, out = ctx.cost_grad_device(sc.T, co, repeats=1)
    ms, out = ctx.cost_grad_device(sc.T, co, repeats=5)
    ctx.executed_evals(True)
    ctx.cost_grad_device(sc.T, co, repeats=1)
    ev = ctx.executed_evals(False)
    t0 = time.time(); c1, gT1, gC1 = ctx.cost_grad(sc.T, co); t1 = time.time()
    print(f"P={P}: {ms:.3f} ms/eval  -> {P/ms*1e3:.3e} pts/s; n_inside={out[-1]}; cost={out[0]:.10g}; lane-evals/pt={ev/P:.1f}; e2e cost_grad {1e3*(t1-t0):.3f} ms")
