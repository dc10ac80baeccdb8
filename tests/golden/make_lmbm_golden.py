"""Generate tests/golden/lmbm_trace_star_400.npz: the iterates visited by the REFERENCE'S OWN outer solver — the prebuilt
LMBM binary src/utils/include/utils/lmbm.so (Fortran 77 + C wrapper, lmbm.cpp) with its default lmbm_parameter_t
(back_end_optimizer.cpp:29) — while it minimises the oracle's costFunctionLmbmParallel on a 400-point star scene.

    python tests/golden/make_lmbm_golden.py        (build container only: needs /root/reference and scipy's libgfortran)

The .so needs libgfortran.so.5; scipy bundles one under a mangled file name, so this script symlinks it into
oracle/_ref/ (git-ignored) and re-executes itself with LD_LIBRARY_PATH pointing there.  tests/test_gpu_parity.py
replays the recorded x_k through svsdf_evaluate on the GPU (no LMBM, no /root/reference needed at run time)."""
import ctypes as C
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
LMBM = "/root/reference/src/utils/include/utils/lmbm.so"

if os.environ.get("SVSDF_LMBM_REEXEC") != "1":
    os.makedirs(REFDIR, exist_ok=True)
    cand = sorted(glob.glob(os.path.join(os.path.dirname(C.__file__), "..", "site-packages", "scipy.libs", "libgfortran-*.so.5*")))
    if not cand:
        import scipy
        cand = sorted(glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libgfortran-*.so.5*")))
    link = os.path.join(REFDIR, "libgfortran.so.5")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(os.path.abspath(cand[0]), link)
    env = dict(os.environ, SVSDF_LMBM_REEXEC="1", LD_LIBRARY_PATH=REFDIR + ":" + os.path.dirname(os.path.abspath(cand[0])) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    os.execve(sys.executable, [sys.executable] + sys.argv, env)

import numpy as np  # noqa: E402

sys.path.insert(0, ROOT)
from implicit_svsdf_planner_b200 import scenes  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


class LmbmParam(C.Structure):  # lmbm.h:15-174 (defaults are the struct's member initialisers)
    _fields_ = [("timeout", C.c_float), ("bundle_size", C.c_int), ("ini_corrections", C.c_int), ("max_corrections", C.c_int),
                ("exponent_distmeasure", C.c_int), ("max_iterations", C.c_int), ("max_evaluations", C.c_int), ("past", C.c_int),
                ("verbose", C.c_int), ("update_method", C.c_int), ("scaling_strategy", C.c_int), ("delta_past", C.c_double),
                ("f_rel_eps", C.c_double), ("f_lower_bound", C.c_double), ("terminate_param1", C.c_double), ("terminate_param2", C.c_double),
                ("distance_measure", C.c_double), ("sufficient_dec", C.c_double), ("max_stepsize", C.c_double)]


def main():
    sc = scenes.make_scene("star", 8, 400, clearance=2.6, seed_map=777)
    orc = O.Oracle("star", weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, threads=O.num_procs())
    orc.set_points(sc.points)
    orc.set_conditions(sc.init_s, sc.final_s, sc.N)
    L = C.CDLL(LMBM)
    opt = getattr(L, "_ZN4lmbm13lmbm_optimizeEiPdS0_PFdPvPKdS0_iES1_PFiS1_S3_iEPNS_16lmbm_parameter_tE")
    EVAL = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)
    PROG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int)
    xs, fs, gs = [], [], []

    def ev(_inst, xp, gp, n):
        x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
        f, g = orc.evaluate(x)
        np.ctypeslib.as_array(gp, shape=(n,))[:] = g
        xs.append(x); fs.append(f); gs.append(g.copy())
        return float(f)

    iters = []

    def prog(_inst, xp, k):
        iters.append(k)
        return 0

    p = LmbmParam(300.0, 2, 7, 15, 2, 10000, 20000, 10, -1, 0, 0, 1.0e-8, 1.0e+4, -1.0e+60, 1.0e-6, 1.0e-6, 0.5, 1.0e-4, 1.5)
    p.max_evaluations = 400  # keep the fixture small; the reference's default is 20000
    x = sc.x0.copy()
    fx = C.c_double()
    opt.restype = C.c_int
    ret = opt(C.c_int(x.size), x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(fx), EVAL(ev), None, PROG(prog), C.byref(p))
    print("LMBM return", ret, "final f", fx.value, "evaluations", len(fs), "iterations", (max(iters) if iters else 0), "f0", fs[0])
    keep = np.unique(np.r_[np.arange(0, len(fs), max(1, len(fs) // 40)), len(fs) - 1])
    np.savez_compressed(os.path.join(HERE, "lmbm_trace_star_400.npz"), shape="star", N=sc.N, points=sc.points, init_s=sc.init_s,
                        final_s=sc.final_s, weight_p=sc.weight_p, safety_hor=sc.safety_hor, rho=sc.rho, x0=sc.x0,
                        xs=np.array(xs)[keep], fs=np.array(fs)[keep], gs=np.array(gs)[keep], all_fs=np.array(fs), lmbm_return=ret,
                        final_f=fx.value, final_x=x, n_evaluations=len(fs))


if __name__ == "__main__":
    main()
