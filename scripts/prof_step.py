"""Profiling driver: a few device-resident cost+grad evaluations (config 2 workload) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from implicit_svsdf_planner_b200 import api, scenes

P = int(os.environ.get("SVSDF_P", "200000"))
N = int(os.environ.get("SVSDF_N", "8"))
shape = os.environ.get("SVSDF_SHAPE", "star")
strict = os.environ.get("SVSDF_STRICT", "1") == "1"
reps = int(os.environ.get("SVSDF_REPS", "3"))
mesh = None
if os.environ.get("SVSDF_MESH", "0") == "1":  # the reference's star.obj through the triangle-mesh functor
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fwn_ref.npz"))
    mesh = (g["star_V"], g["star_F"])
sc = scenes.make_scene(shape, N, P)
ctx = api.Context(shape, strict_fp=strict, mesh=mesh)
ctx.set_points(sc.points)
co = sc.coeffs_colmajor()
for _ in range(reps):
    ms, out = ctx.cost_grad_device(sc.T, co, repeats=1)
print(f"shape={shape} N={N} P={P} strict={strict}: {ms:.3f} ms/eval, cost {out[0]:.10g}, n_inside {out[-1]:.0f}")
