"""Scratch timing of the stage entry points on one frame (developer tool)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle import pyoracle as O
from perception_b200 import api, synth
from perception_b200.params import default_params

p = default_params("multi8")
d = synth.depth_frame("multi8", 0)
pts = O.unproject(d, p.fx, p.fy, p.cx, p.cy, p.depth_scale)
pz, _ = O.passthrough(pts, 2, p.pass_z_min, p.pass_z_max)
px, _ = O.passthrough(pz, 0, p.pass_x_min, p.pass_x_max)
vg = O.voxel_grid(px, p.leaf)
sac = O.sac_plane(vg["vox"])
rem, _ = O.extract(vg["vox"], sac["inliers"], True)
print("remain", len(rem))
with api.CuboidCuda(p, max_points=640 * 480, max_batch=1) as cc:
    for n in (1400, 3000, 6000, len(rem)):
        sub = rem[:n]
        cc.cluster(sub)
        t = time.perf_counter()
        for _ in range(5):
            idx, off = cc.cluster(sub)
        dt = (time.perf_counter() - t) / 5
        print(n, "clusters", len(off) - 1, "ms per call %.3f" % (dt * 1e3))
    t = time.perf_counter()
    for _ in range(5):
        cc.segment_plane(vg["vox"])
    print("segment_plane ms %.3f" % ((time.perf_counter() - t) / 5 * 1e3))
