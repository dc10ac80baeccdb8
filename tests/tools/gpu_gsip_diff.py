"""How far are the interior-branch (GSIP) per-point results of the strict build from the oracle's?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from implicit_svsdf_planner_b200 import api, scenes
from oracle import oracle_py as O

for sc in (scenes.make_scene("star", 8, 400, clearance=2.35), scenes.make_scene("star", 8, 3000, clearance=2.5), scenes.make_scene("sdHorseshoe", 8, 2000, clearance=1.9)):
    co = sc.coeffs_colmajor()
    opt = api.TrajOptimizer(sc.shape, strict_fp=True)
    sv = opt.sv_manager
    sv.updateTraj(sc.T, co)
    orc = O.Oracle(sc.shape, threads=O.num_procs())
    orc.set_traj(sc.T, co)
    p = np.c_[sc.points[:, :2], np.zeros(sc.P)]
    s_c, t_c, g_c, r_c = orc.query(p)
    s_g, t_g, g_g, r_g = sv.getTrueSDFofSweptVolume(p)
    ins = r_c > 0
    print(sc.shape, "inside", int(ins.sum()), "rounds equal", bool(np.array_equal(r_c, r_g)),
          "max|dsdf|", np.abs(s_g - s_c)[ins].max() if ins.any() else 0, "max|dt|", np.abs(t_g - t_c)[ins].max() if ins.any() else 0,
          "max|dg|", np.abs(g_g - g_c)[ins].max() if ins.any() else 0,
          "n bitwise equal sdf", int((s_g == s_c)[ins].sum()), "t", int((t_g == t_c)[ins].sum()), "g", int((g_g == g_c).all(axis=1)[ins].sum()))

# callback level: f and g of svsdf_evaluate vs the oracle on the LMBM golden scene
sc = scenes.make_scene("star", 8, 400, clearance=2.6, seed_map=777)
opt = api.TrajOptimizer("star", weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, strict_fp=True)
opt.parallel_points = sc.points
opt.setConditions(sc.init_s, sc.final_s, sc.N)
orc = O.Oracle("star", weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, threads=O.num_procs())
orc.set_points(sc.points)
orc.set_conditions(sc.init_s, sc.final_s, sc.N)
f0, g0 = orc.evaluate(sc.x0)
f1, g1 = opt.costFunction(sc.x0)
print("evaluate: rel df", abs(f1 - f0) / abs(f0), "normwise dg", np.linalg.norm(g1 - g0) / np.linalg.norm(g0), "max|dg|", np.abs(g1 - g0).max(), "max|g|", np.abs(g0).max())
print("dg per entry", (g1 - g0))
co = sc.coeffs_colmajor()
c0, gT0, gC0, _, ins = orc.cost_grad(sc.T, co)
c1, gT1, gC1 = opt.addSaftyPenaOnSweptVolumeParallelTrueSDF(sc.T, co)
print("cost_grad: rel dc", abs(c1 - c0) / abs(c0), "gC", np.linalg.norm(gC1 - gC0) / np.linalg.norm(gC0), "gT", np.abs(gT1 - gT0).max(), np.abs(gT0).max(), "inside", ins)
