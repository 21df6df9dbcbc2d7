"""Developer tool (GPU box): single-frame latency of cuboid_process_cloud / cuboid_process_batch(1) (p50 over N calls)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from perception_b200 import api, pcd, synth
from perception_b200.params import default_params
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
p = default_params("cuboid")
depth = synth.depth_frame("cuboid1", 0)
with api.CuboidCuda(p, max_points=depth.size, max_batch=1) as cc:
    cc.set_template(0, pcd.template_points(0.2, 0.1, 0.03, 0.002))
    cc.set_option(api.OPT_TAPS, 0)
    cloud = np.ascontiguousarray(cc.unproject(depth))
    import torch
    pinned = torch.from_numpy(cloud).pin_memory().numpy()
    for name, fn in (("process_cloud", lambda: cc.process_cloud(cloud)), ("process_cloud_pinned", lambda: cc.process_cloud(pinned)),
                     ("process_batch_1", lambda: cc.process_batch(depth[None]))):
        for _ in range(10):
            r = fn()
        ts = []
        for _ in range(calls):
            t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
        ts = np.asarray(ts) * 1e3
        r0 = r[0] if name == "process_batch_1" else r
        print(name, "p50 %.3f ms p99 %.3f ms" % (np.percentile(ts, 50), np.percentile(ts, 99)), "iters", r0.cluster[0].iterations, "n_remain", r0.n_remain)
