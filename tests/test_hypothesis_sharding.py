"""Hypothesis-sharded low-latency mode (SURVEY.md 8e second axis; north_star: "ICP initial-pose hypotheses are partitioned across
the GPUs ... NCCL ... best-fitness/pose reduction"): the initial-pose hypotheses of a cluster are split over ranks, every rank
runs the same frame, one all-gather of 80-byte records, exact (fitness, guess id) arg-min. The winner and its pose bytes must
equal the run that holds all hypotheses on one GPU, for any split."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402
from perception_b200 import api, synth  # noqa: E402
from perception_b200 import dist as pd  # noqa: E402
from perception_b200.params import ClusterResult, default_params  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def bench_rotations():
    sys.path.insert(0, ROOT)
    import bench
    return bench.guess_rotations()


def _oracle_cluster_source(depth, p):
    pts = O.unproject(depth, p.fx, p.fy, p.cx, p.cy, p.depth_scale)
    a, _ = O.passthrough(pts, 2, p.pass_z_min, p.pass_z_max)
    a, _ = O.passthrough(a, 0, p.pass_x_min, p.pass_x_max)
    v = O.voxel_grid(a, p.leaf)["vox"]
    s = O.sac_plane(v, p.sac_threshold, p.sac_max_iter, p.sac_prob, p.sac_seed, 1)
    rem, _ = O.extract(v, s["inliers"], True)
    idx, off = O.cluster(rem, p.cluster_tol, p.cluster_min, p.cluster_max)
    return rem[idx[off[0]:off[1]]]


def test_record_layout_and_exact_tie_rule():
    assert C.sizeof(api.GuessRecord) == 80
    a, b = ClusterResult(), ClusterResult()
    a.fitness, a.best_guess, a.iterations, a.state, a.converged = 1.25e-6, 40, 77, 4, 1
    # differs from a only below the 2^-36 the legacy key keeps: the exact rule must still prefer the smaller fitness
    b.fitness, b.best_guess, b.iterations, b.state, b.converged = float(np.nextafter(1.25e-6, 0.0)), 50, 12, 2, 1
    for k in range(16):
        a.T[k], b.T[k] = float(k), float(-k)
    out = ClusterResult()
    out.size = 321
    api.reduce_guess_records([api.guess_record(a), api.guess_record(b)], 4e-4, out)
    assert out.best_guess == 50 and out.iterations == 12 and out.state == 2 and out.converged == 1 and out.accepted == 1
    assert out.size == 321 and list(out.T) == [float(-k) for k in range(16)]
    assert api.pack_fitness_key(a.fitness, 40) < api.pack_fitness_key(b.fitness, 50)   # the legacy key gets this one wrong
    b.fitness = a.fitness                                                             # exact tie -> lowest guess id
    api.reduce_guess_records([api.guess_record(b), api.guess_record(a)], 4e-4, out)
    assert out.best_guess == 40
    a.fitness = float("nan")
    api.reduce_guess_records([api.guess_record(a), api.guess_record(b)], 4e-4, out)
    assert out.best_guess == 50                                                       # NaN never wins


def _cpu_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # REAL ICP outputs (the CPU oracle's) for 8 hypotheses about the cluster centroid, split over the ranks
        p = default_params("cuboid")
        p.icp_max_iter = 30
        from perception_b200 import pcd
        tm = pcd.load_pcd(os.path.join(ROOT, "tests", "golden", "template_cuboid_L200_W100_H30_3faces.pcd"))
        src = _oracle_cluster_source(synth.depth_frame("bench", 7), p)
        rots = bench_rotations()[:8]
        g0, cnt = pd.shard_range(len(rots), rank, world)
        best = None
        for g in range(g0, g0 + cnt):
            G = O.guess_about_centroid(src, rots[g])
            r = O.icp(src, tm, guess=G, max_iter=p.icp_max_iter, rel_mse=p.icp_rel_mse)
            if best is None or r["fitness"] < best[1]["fitness"]:
                best = (g, r)
        fr = api.FrameResult()
        fr.n_clusters = 1
        c = fr.cluster[0]
        c.size, c.best_guess, c.fitness, c.iterations, c.state, c.converged = len(src), best[0], best[1]["fitness"], best[1]["iters"], best[1]["state"], best[1]["converged"]
        for k in range(16):
            c.T[k] = float(best[1]["T"].reshape(16)[k])
        res = (api.FrameResult * 1)(fr)
        pd.reduce_hypotheses(res, p.icp_fitness_gate)
        w = res[0].cluster[0]
        q.put((rank, w.best_guess, w.fitness, w.iterations, bytes(bytearray(np.asarray(list(w.T), np.float32).tobytes()))))
    finally:
        dist.destroy_process_group()


def test_reduction_of_real_icp_outputs_names_the_single_process_winner():
    """gloo, world size 2, CPU only: the records come from the oracle's ICP (real poses and fitness values)."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = sorted(q.get(timeout=300) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = default_params("cuboid")
    p.icp_max_iter = 30
    from perception_b200 import pcd
    tm = pcd.load_pcd(os.path.join(ROOT, "tests", "golden", "template_cuboid_L200_W100_H30_3faces.pcd"))
    src = _oracle_cluster_source(synth.depth_frame("bench", 7), p)
    rots = bench_rotations()[:8]
    runs = [O.icp(src, tm, guess=O.guess_about_centroid(src, rots[g]), max_iter=p.icp_max_iter, rel_mse=p.icp_rel_mse) for g in range(8)]
    want = min(range(8), key=lambda g: (runs[g]["fitness"], g))
    for rank, bg, fit, iters, Tb in outs:
        assert bg == want and fit == runs[want]["fitness"] and iters == runs[want]["iters"]
        assert Tb == np.asarray(runs[want]["T"], np.float32).tobytes()


def _gpu_worker(rank, world, port, backend, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if backend == "nccl" else 0          # gloo: both ranks share cuda:0 (one-GPU boxes)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from perception_b200 import pcd
        p = default_params("cuboid")
        rots = bench_rotations()
        p.n_guess, p.guess_mode = len(rots), 1
        tm = pcd.load_pcd(os.path.join(ROOT, "tests", "golden", "template_cuboid_L200_W100_H30_3faces.pcd"))
        depth = synth.depth_batch("bench", [3, 4])
        g0, cnt = pd.shard_range(len(rots), rank, world)
        with api.CuboidCuda(p, device=dev, max_points=640 * 480, max_batch=2) as h:
            h.set_template(0, tm)
            h.set_guesses(rots[g0:g0 + cnt], mode=1)
            h.set_guess_offset(g0)
            res = h.process_batch(depth)
        pd.reduce_hypotheses(res, p.icp_fitness_gate)
        q.put((rank, [bytes(r) for r in res]))
    finally:
        dist.destroy_process_group()


def _run_gpu_world(world, backend):
    import torch.multiprocessing as mp
    from perception_b200 import pcd
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = sorted(q.get(timeout=600) for _ in range(world))
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    p = default_params("cuboid")
    rots = bench_rotations()
    p.n_guess, p.guess_mode = len(rots), 1
    tm = pcd.load_pcd(os.path.join(ROOT, "tests", "golden", "template_cuboid_L200_W100_H30_3faces.pcd"))
    depth = synth.depth_batch("bench", [3, 4])
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=2) as h:
        h.set_template(0, tm)
        h.set_guesses(rots, mode=1)
        h.set_option(api.OPT_TAPS, 1)
        one = h.process_batch(depth)
    for r in one:                                   # the per-rank parity tap is not carried through the reduction
        for c in range(api.MAX_CLUSTERS):
            r.cluster[c].corr_hash = 0
    want = [bytes(r) for r in one]
    for rank, got in outs:
        assert got == want, "rank %d of %d (%s) disagrees with the single-GPU run" % (rank, world, backend)
    return one


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_hypotheses_split_over_ranks_equal_the_single_gpu_run_gloo(world):
    """Real cuboid_process_batch outputs, 64 hypotheses split 2- and 4-way (ranks share cuda:0, records travel over gloo)."""
    one = _run_gpu_world(world, "gloo")
    assert one[0].n_clusters == 1 and 0 <= one[0].cluster[0].best_guess < 64


@pytest.mark.gpu
def test_hypotheses_split_over_gpus_equal_the_single_gpu_run_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    _run_gpu_world(2, "nccl")
    if n >= 8:
        _run_gpu_world(8, "nccl")
