"""Developer probe: how many source-template pairs the ICP kernel evaluates as a function of the iteration cap."""
import sys
sys.path.insert(0, ".")
import numpy as np
from perception_b200 import api, pcd, synth
from perception_b200.params import default_params

tm = pcd.template_points(0.2, 0.1, 0.03, 0.002)
depth = synth.depth_batch("bench", [0])
prev = 0
for it in (1, 2, 3, 5, 10, 20, 40, 80):
    p = default_params("cuboid")
    p.icp_max_iter = it
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=1) as cc:
        cc.set_template(0, tm)
        r = cc.process_batch(depth)
        ev, br = cc.icp_work()
    print("max_iter %3d iters %3d evaluated pairs %10d (%.4f of brute force %d) per-pass-avg %.0f" % (it, r[0].cluster[0].iterations, ev, ev / br, br, ev / (r[0].cluster[0].iterations + 1)))
