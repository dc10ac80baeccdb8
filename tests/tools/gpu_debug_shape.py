import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from implicit_svsdf_planner_b200 import api, scenes
from oracle import oracle_py as O
shape, N = sys.argv[1], int(sys.argv[2])
sc = scenes.make_scene(shape if shape in scenes.START_GOAL else "star", N, 600, clearance=2.0)
co = sc.coeffs_colmajor()
ctx = api.Context(shape, strict_fp=True); ctx.set_points(sc.points)
orc = O.Oracle(shape, threads=O.num_procs()); orc.set_points(sc.points)
c0, gT0, gC0, pp, inside = orc.cost_grad(sc.T, co, per_point=True)
c1, gT1, gC1 = ctx.cost_grad(sc.T, co)
rel = lambda a,b: np.linalg.norm(a-b)/np.linalg.norm(b)
print("cost", c0, c1, "gC rel", rel(gC1,gC0), "gT abs", np.abs(gT1-gT0).max())
p = np.c_[sc.points[:,:2], np.zeros(sc.P)]
orc.set_traj(sc.T, co)
s_c, t_c, g_c, r_c = orc.query(p)
s_g, t_g, g_g, r_g = ctx.query(sc.T, co, p)
act = pp[:,6] > 0
d = np.abs(t_g - t_c)
idx = np.argsort(-d)[:8]
for i in idx: print(i, "act", act[i], "dt %.3e"%d[i], "t_c %.9f t_g %.9f"%(t_c[i], t_g[i]), "sdf_c %.12f sdf_g %.12f"%(s_c[i], s_g[i]), "g_c", g_c[i,:2], "g_g", g_g[i,:2])
dg = np.abs(g_g-g_c).max(axis=1)
idx = np.argsort(-dg)[:8]
print("largest grad diffs")
for i in idx: print(i, "act", act[i], "dg %.3e"%dg[i], "dt %.3e"%d[i], "t_c %.9f"%t_c[i], "sdf %.9f"%s_c[i], g_c[i,:2], g_g[i,:2])
