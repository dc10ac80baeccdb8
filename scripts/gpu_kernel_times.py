"""Quick device-time breakdown (CUDA events) for configs 1-3, no CPU work."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from implicit_svsdf_planner_b200 import api, scenes
for (shape, N, P, cl) in (("star", 8, 2000, 2.75), ("star", 8, 200_000, 2.75), ("sdHorseshoe", 16, 500_000, 2.15), ("star", 8, 400, 2.35)):
    sc = scenes.make_scene(shape, N, P, clearance=cl)
    ctx = api.Context(shape); ctx.set_points(sc.points); co = sc.coeffs_colmajor()
    ctx.cost_grad_device(sc.T, co, repeats=3, fetch=False)
    ms, out = ctx.cost_grad_device(sc.T, co, repeats=10)
    t0 = time.perf_counter()
    for _ in range(20): ctx.cost_grad(sc.T, co)
    e2e = (time.perf_counter() - t0) / 20 * 1e3
    print(f"{shape} N={N} P={P}: {ms:.4f} ms/eval device; e2e svsdf_cost_grad {e2e:.4f} ms; kernels {[round(v,4) for v in ctx.last_kernel_ms()]} n_inside {out[-1]:.0f}", flush=True)
