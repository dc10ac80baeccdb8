"""Runs the BASELINE.json configurations on one GPU and prints one JSON line per config (used to fill BASELINE.md §3).

  config 1: star, N=8, 2k points        (parity scene; GPU vs CPU oracle on all points)
  config 2: star, N=8, 200k points      (bench.py's workload; here: cost+grad and a full L-BFGS run, GPU and CPU)
  config 3: sdHorseshoe, N=16, 500k points
  config 4m: the reference's shapes/star.obj (152 v / 300 f, tests/golden/fwn_ref.npz) through the triangle-mesh functor
            (BasicShape::getonlySDF_igl: the reference's float winding-number hierarchy reproduced bit for bit + exact closest triangle), N=16, 500k
  config 4: mesh shape: outline of the reference's shapes/star.obj (40 vertices) through the Polygon fallback functor (what this
            release of the reference uses for non-analytic shapes; the libigl path is unreachable there), N=16, 500k
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from implicit_svsdf_planner_b200 import api, scenes
from oracle import oracle_py as O

def nrel(a, b): return float(np.linalg.norm(np.asarray(a) - b) / max(np.linalg.norm(b), 1e-300))

def star_obj_polygon():
    """2-D outline of the reference's shapes/star.obj (tests/golden/obj_outlines.json), ordered by angle: the mesh shape of
    config 4 handed to the Polygon fallback functor (what this release of the reference uses for non-analytic shapes)."""
    tests_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xy = np.array(json.load(open(os.path.join(tests_dir, "golden", "obj_outlines.json")))["star"]["outline_xy"])
    return xy[np.argsort(np.arctan2(xy[:, 1], xy[:, 0]))].reshape(-1)


def run(name, shape, N, P, clearance, cpu_points, lbfgs_iters=60, scene_shape=None, polygon=None, mesh=None):
    sc = scenes.make_scene(scene_shape or shape, N, P, clearance=clearance)
    co = sc.coeffs_colmajor()
    ctx = api.Context(shape, weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, polygon=polygon, mesh=mesh)
    ctx.set_points(sc.points)
    ctx.cost_grad_device(sc.T, co, repeats=2, fetch=False)
    ms, out = ctx.cost_grad_device(sc.T, co, repeats=5)
    km = ctx.last_kernel_ms()
    ctx.executed_evals(True); ctx.cost_grad_device(sc.T, co, repeats=1, fetch=False); ev = ctx.executed_evals(False)
    # CPU oracle (glibc variant = reference behaviour) on a strided sample, best thread count
    stride = max(1, P // cpu_points)
    sub = sc.points[::stride]
    nproc = O.num_procs()
    best = None
    for th in sorted({nproc, max(1, nproc // 2), max(1, nproc // 4), int(1.5 * nproc)}):
        orc = O.Oracle(shape, weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, threads=th, variant="glibc", polygon=polygon, mesh=mesh)
        orc.set_points(sub)
        sec, _ = orc.time_cost_grad(sc.T, co, warm=1, reps=2)
        if best is None or sec < best[0]: best = (sec, th)
    cpu_pts_s = sub.shape[0] / best[0]
    # parity on the sample: strict GPU vs default oracle (bitwise per point), cost/grad
    orc = O.Oracle(shape, weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, threads=best[1], polygon=polygon, mesh=mesh)
    orc.set_points(sub)
    c0, gT0, gC0, _, inside = orc.cost_grad(sc.T, co)
    ctx2 = api.Context(shape, weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, polygon=polygon, mesh=mesh)
    ctx2.set_points(sub)
    c1, gT1, gC1 = ctx2.cost_grad(sc.T, co)
    rec = dict(config=name, shape=shape, N=N, P=P, gpu_ms_per_eval=ms, gpu_pts_per_s=P / ms * 1e3, kernel_ms=km,
               lane_evals_per_point=ev / P, n_inside=int(out[-1]), cpu_pts_per_s=cpu_pts_s, cpu_threads=best[1], cpu_cores=nproc,
               cpu_sample_points=int(sub.shape[0]), speedup=P / ms * 1e3 / cpu_pts_s,
               rel_err_cost=abs(c1 - c0) / abs(c0), rel_err_gradC=nrel(gC1, gC0),
               rel_err_gradT_budget=float(np.linalg.norm(gT1 - gT0) / (np.linalg.norm(gT0) + 1e-3 * np.linalg.norm(gC0))))
    if lbfgs_iters:
        params = api.default_lbfgs_params(mem_size=16, past=3, delta=1e-6, g_epsilon=0.0, max_iterations=lbfgs_iters, min_step=1e-32)
        rc, x, T, b, st = ctx.optimize(sc.init_s, sc.final_s, sc.x0, sc.N, params)
        rec.update(lbfgs=dict(iterations=st["iterations"], evaluations=st["evaluations"], status=st["status"], seconds=st["seconds"],
                              iters_per_s=st["iterations"] / st["seconds"], evals_per_s=st["evaluations"] / st["seconds"], final_cost=st["final_cost"]))
        # CPU: a few L-BFGS iterations on the same sample to get iters/s of the reference path
        orc.set_conditions(sc.init_s, sc.final_s, sc.N)
        t0 = time.time(); ret, xo, so = orc.lbfgs(sc.x0, mem_size=16, past=3, delta=1e-6, max_iterations=3); dt = time.time() - t0
        rec.update(cpu_lbfgs=dict(iterations=so["iters"], evaluations=so["evals"], seconds=dt, evals_per_s_scaled_to_full_P=so["evals"] / dt / stride,
                                  note=f"oracle on {sub.shape[0]} points; rate scaled by 1/{stride} to the full point count"))
    print(json.dumps(rec), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "3", "4"]
    if "1" in which: run("1", "star", 8, 2000, 2.75, 2000)
    if "2" in which: run("2", "star", 8, 200_000, 2.75, 50_000)
    if "3" in which: run("3", "sdHorseshoe", 16, 500_000, 2.15, 50_000)
    if "4m" in which:  # the same mesh through the triangle-mesh functor (getonlySDF_igl)
        g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "fwn_ref.npz"))
        P4 = int(os.environ.get("SVSDF_MESH_P", "500000"))
        run("4m", "star_obj_mesh_sdf", 16, P4, 2.75, 3000, lbfgs_iters=int(os.environ.get("SVSDF_MESH_LBFGS", "4")), scene_shape="sdHorseshoe",
            mesh=(g["star_V"], g["star_F"]))
    if "4" in which: run("4", "star_obj_outline_polygon", 16, 500_000, 2.75, 25_000, scene_shape="sdHorseshoe", polygon=star_obj_polygon())
