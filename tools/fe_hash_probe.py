"""Developer tool (GPU box): which path k_frontend takes per frame (CUBOID_FE_HASH=2 marks it in status bits 8..13) and stage times."""
import collections, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["CUBOID_FE_HASH"] = "2"
from perception_b200 import api, pcd, synth
from perception_b200.params import default_params
p = default_params("cuboid")
n = 256
depth = synth.depth_batch("bench", range(n))
with api.CuboidCuda(p, max_points=640 * 480, max_batch=n) as cc:
    cc.set_template(0, pcd.template_points(0.2, 0.1, 0.03, 0.002))
    cc.set_option(api.OPT_TAPS, 0)
    res = cc.process_batch(depth)
c = collections.Counter((r.status >> 8) & 63 for r in res)
print("path marks (32 = hash path done; 1 bits, 2 table/abort, 4 long run, 8 count mismatch, 16 sum mismatch):", dict(c))
print("V min/mean/max", min(r.n_voxels for r in res), np.mean([r.n_voxels for r in res]), max(r.n_voxels for r in res))
