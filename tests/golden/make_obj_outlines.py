"""Generate tests/golden/obj_outlines.json from the reference's own shape meshes.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_obj_outlines.py

For every src/plan_manager/shapes/<name>.obj (Meshlab exports of the 2-D shapes extruded to thin slabs,
SURVEY.md A.12) it records the outline of the top face: the vertices of top-face edges that belong to exactly one
top-face triangle, plus the top-face area.  tests/test_oracle_shapes.py checks that the restated analytic SDFs
vanish on these outlines and enclose the same area — the only pin on the shape functors that comes from the
reference's own fixtures.
"""
import collections
import glob
import json
import os

import numpy as np

REF = "/root/reference/src/plan_manager/shapes"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "obj_outlines.json")


def load(path):
    V, F = [], []
    for line in open(path):
        if line.startswith("v "):
            V.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            F.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
    return np.array(V), np.array(F)


def main():
    out = {}
    for f in sorted(glob.glob(REF + "/*.obj")):
        name = os.path.basename(f)[:-4]
        V, F = load(f)
        top = [fc for fc in F if (V[fc, 2] > 0).all()]
        cnt = collections.Counter()
        for a, b, c in top:
            for e in ((a, b), (b, c), (c, a)):
                cnt[tuple(sorted(e))] += 1
        bv = sorted({v for e, n in cnt.items() if n == 1 for v in e})
        area = 0.0
        for a, b, c in top:
            u, w = V[b, :2] - V[a, :2], V[c, :2] - V[a, :2]
            area += 0.5 * abs(u[0] * w[1] - u[1] * w[0])
        out[name] = {
            "source": f"src/plan_manager/shapes/{name}.obj",
            "n_vertices": int(V.shape[0]),
            "n_faces": int(F.shape[0]),
            "top_area": area,
            "outline_xy": [[float(V[i, 0]), float(V[i, 1])] for i in bv],
            "all_top_xy": [[float(x), float(y)] for x, y, z in V if z > 0],
        }
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", OUT, {k: len(v["outline_xy"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
