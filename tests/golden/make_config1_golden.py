"""Generate tests/golden/config1_star_2k.npz and config_inside_400.npz: golden vectors of the CPU oracle for
BASELINE config 1 (star, 8-piece MINCO, 2 000 query points, one cost+gradient evaluation) and for a small scene
with ~6 % of the points inside the swept volume (GSIP branch).

    python tests/golden/make_config1_golden.py

The reference has no golden vectors for this path, so these are the oracle's own outputs at
the commit that created them; tests/test_golden.py (CPU) checks the oracle still reproduces them and
tests/test_gpu_parity.py (GPU) checks the CUDA path against them without needing the oracle at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from implicit_svsdf_planner_b200 import scenes  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, scene):
    co = scene.coeffs_colmajor()
    orc = O.Oracle(scene.shape, weight_p=scene.weight_p, safety_hor=scene.safety_hor, rho=scene.rho, threads=O.num_procs())
    orc.set_points(scene.points)
    cost, gT, gC, pp, inside = orc.cost_grad(scene.T, co, per_point=True)
    orc.set_traj(scene.T, co)
    pts0 = np.c_[scene.points[:, :2], np.zeros(scene.P)]
    q_sdf, q_ts, q_g, q_rounds = orc.query(pts0)
    orc.set_conditions(scene.init_s, scene.final_s, scene.N)
    f, g = orc.evaluate(scene.x0)
    np.savez_compressed(
        os.path.join(HERE, name),
        shape=scene.shape, N=scene.N, T=scene.T, coeffs_colmajor=co, points=scene.points, init_s=scene.init_s,
        final_s=scene.final_s, q=scene.q, x0=scene.x0, weight_p=scene.weight_p, safety_hor=scene.safety_hor,
        rho=scene.rho, cost=cost, gradT=gT, gradC=gC, per_point=pp, n_inside=inside, query_sdf=q_sdf,
        query_tstar=q_ts, query_grad=q_g, query_rounds=q_rounds, eval_f=f, eval_g=g,
    )
    print(name, "cost", cost, "inside", inside, "f", f)


if __name__ == "__main__":
    make("config1_star_2k.npz", scenes.make_scene("star", 8, 2000))
    make("config_inside_400.npz", scenes.make_scene("star", 8, 400, clearance=2.35))
