"""GPU (-m gpu): the header-only C++ mirror of the reference call surface (include/svsdf.hpp), compiled with g++ against
libsvsdf_b200.so and driven like the reference's plan_manager, reproduces the golden vectors."""
import os
import subprocess

import numpy as np
import pytest

from implicit_svsdf_planner_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_matches_golden(tmp_path):
    G = np.load(os.path.join(ROOT, "tests", "golden", "config1_star_2k.npz"))
    exe = tmp_path / "mirror_main"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_main.cpp"), "-o", str(exe), "-L", libdir, "-lsvsdf_b200",
                           f"-Wl,-rpath,{libdir}"])
    N, P = int(G["N"]), G["points"].shape[0]
    inp = tmp_path / "in.txt"
    with open(inp, "w") as f:
        f.write(f"{str(G['shape'])} {N} {P} {float(G['weight_p'])!r} {float(G['safety_hor'])!r} {float(G['rho'])!r}\n")
        for arr in (G["T"], G["coeffs_colmajor"], G["init_s"].T.reshape(-1), G["final_s"].T.reshape(-1), G["x0"]):
            f.write(" ".join(repr(float(v)) for v in np.asarray(arr).reshape(-1)) + "\n")
        for p in G["points"]:
            f.write(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
    out = subprocess.run([str(exe), str(inp)], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    rows = [np.array([float(v) for v in ln.split()]) for ln in out]
    assert rows[0][0] == 0  # registry id of "star"
    assert abs(rows[1][0] - 1.0) < 1e-12  # star sdf 1 m above the tip
    assert np.allclose(rows[2], [0, 1, 0], atol=1e-6)
    for i in range(5):
        q = rows[3 + i]
        if G["query_rounds"][i] == 0:
            assert q[0] == G["query_sdf"][i] and q[1] == G["query_tstar"][i] and q[2] == G["query_grad"][i, 0] and q[3] == G["query_grad"][i, 1]
    rc, cost = rows[8]
    assert rc == 0 and abs(cost - (1.5 + float(G["cost"]))) <= 1e-12 * cost
    assert np.abs(rows[9] - (0.25 + G["gradT"])).max() <= 1e-9 * max(1.0, np.abs(G["gradT"]).max())
    assert np.abs(rows[10] - (-0.5 + G["gradC"])).max() <= 1e-9 * np.abs(G["gradC"]).max()
    f, cpos, cother, ctot = rows[11]
    assert abs(f - float(G["eval_f"])) <= 1e-12 * f and abs(ctot - f) <= 1e-12 * f and abs(cpos - float(G["cost"])) <= 1e-12 * cpos
    assert np.linalg.norm(rows[12] - G["eval_g"]) <= 1e-8 * np.linalg.norm(G["eval_g"])
    ret, iters, evals, fin = rows[13]
    assert iters >= 1 and evals >= iters and fin < f
