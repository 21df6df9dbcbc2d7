"""Developer tool (GPU box, library built with CUBOID_NVCC_DEFINES=-DCUBOID_ICP_STATS): work-list statistics of the queued ICP search."""
import sys
import numpy as np
sys.path.insert(0, ".")
from perception_b200 import api, pcd, synth
from perception_b200.params import default_params

p = default_params("cuboid")
tm = pcd.template_points(0.2, 0.1, 0.03, 0.002)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
depth = synth.depth_batch("bench", range(n))
import time
with api.CuboidCuda(p, max_points=640 * 480, max_batch=n) as cc:
    t0 = time.time(); cc.set_template(0, tm); print('set_template %.3f s' % (time.time() - t0))
    cc.set_option(api.OPT_TAPS, 0)
    cc.debug_counters()
    cc.process_batch(depth)
    c = cc.debug_counters()
tasks = max(c[0], 1)
print("warp tasks %d  with work list %.3f  fallback %.4f" % (c[0], c[1] / tasks, c[2] / tasks))
print("per task: node rounds %.2f  node items %.1f  leaf rounds %.2f  leaf items %.1f  initial items %.1f" %
      (c[3] / tasks, c[4] / tasks, c[5] / tasks, c[6] / tasks, c[7] / tasks))
print("node-round histogram (tasks that did not fall back):", [round(x / tasks, 3) for x in c[8:30]])
print("candidate table: %d of %d queries answered (%.3f)" % (c[30], c[31], c[30] / max(c[31], 1)))
