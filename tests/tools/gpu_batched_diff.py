"""Diagnostic (GPU box): batched-schedule query vs the oracle on a small scene; prints the first mismatching points."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("SVSDF_FORCE_GRID_OUTER", "2")
os.environ.setdefault("SVSDF_FORCE_BATCHED", "1")
from implicit_svsdf_planner_b200 import api, scenes  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "star"
sc = scenes.make_scene("star", 8, 900, clearance=1.6, seed_map=777)
co = sc.coeffs_colmajor()
p = np.c_[sc.points[:, :2], np.zeros(sc.P)]
ctx = api.Context(shape)
orc = O.Oracle(shape, threads=O.num_procs())
orc.set_traj(sc.T, co)
s_c, t_c, g_c, r_c = orc.query(p)
s_g, t_g, g_g, r_g = ctx.query(sc.T, co, p)
out = r_c == 0
bad = np.flatnonzero(out & ((s_g != s_c) | (t_g != t_c) | (g_g != g_c).any(axis=1)))
print(shape, "points", sc.P, "outside", int(out.sum()), "mismatching", bad.size, "rounds equal", bool(np.array_equal(r_g, r_c)))
for i in bad[:12]:
    print(i, "gpu", s_g[i], t_g[i], g_g[i], "cpu", s_c[i], t_c[i], g_c[i])
print("executed lane-evals per point:", ctx.executed_evals() / max(sc.P, 1))
