"""Fixture for the query-point construction (K3) from the reference's own scene.

Runs in the build container only (needs /root/reference):  python tests/golden/make_k3_golden.py
Copies the 210 points of src/plan_manager/pcds/map_star.pcd (ascii, float32), the start / end of pcds/trajectory_star.txt and the map
parameters of config/star.yaml (occupancy_resolution, sta_threshold, kernel_size) and traj_parlength (plan_manager.cpp:75) into tests/golden/map_star_pcd.npz.
The known answer it carries is the number of occupied voxels, 148 (counted from the same file).
tests/test_oracle_k3.py builds the map with oracle/k3_points.py and checks that count; tests/test_gpu_extract.py runs the device
kernels on the same map.
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/plan_manager"


def read_pcd_ascii(path):
    lines = open(path).read().split("\n")
    i = next(k for k, l in enumerate(lines) if l.startswith("DATA"))
    assert lines[i].strip() == "DATA ascii"
    return np.array([[float(v) for v in l.split()] for l in lines[i + 1:] if l.strip()], dtype=np.float32)


def yaml_value(text, key):
    m = re.search(r"^%s:\s*([-0-9.eE]+)" % re.escape(key), text, re.M)
    return float(m.group(1))


def main():
    pts = read_pcd_ascii(os.path.join(REF, "pcds", "map_star.pcd"))
    traj = open(os.path.join(REF, "pcds", "trajectory_star.txt")).read()
    start = [float(v) for v in re.search(r"Start:\s*(.*)", traj).group(1).split()]
    end = [float(v) for v in re.search(r"End:\s*(.*)", traj).group(1).split()]
    y = open(os.path.join(REF, "config", "star.yaml")).read()
    out = dict(points=pts, start=np.array(start), end=np.array(end), occupancy_resolution=yaml_value(y, "occupancy_resolution"),
               sta_threshold=int(yaml_value(y, "sta_threshold")), kernel_size=int(yaml_value(y, "kernel_size")),
               traj_parlength=3.0,  # plan_manager.cpp:75 (a constant, not in the yaml)
               occupied_voxels=148)
    np.savez_compressed(os.path.join(HERE, "map_star_pcd.npz"), **out)
    for k, v in out.items():
        print(k, np.asarray(v).shape, np.asarray(v).dtype, v if np.asarray(v).size < 4 else "")


if __name__ == "__main__":
    main()
