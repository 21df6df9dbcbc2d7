"""World-size-2 gloo tests of the multi-GPU host logic (frame sharding + best-hypothesis reduction), on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from perception_b200 import api
from perception_b200 import dist as pd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                s, c = pd.shard_range(n, r, world)
                got += list(range(s, s + c))
            assert got == list(range(n))
            counts = [pd.shard_range(n, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def test_pack_key_matches_c_abi():
    rng = np.random.default_rng(0)
    f = np.abs(rng.normal(size=50)) * 1e-5
    g = rng.integers(0, 64, 50)
    k = pd.pack_key(f, g)
    for i in range(50):
        assert int(k[i]) == api.pack_fitness_key(float(f[i]), int(g[i]))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 6 clusters x 8 hypotheses, hypotheses sharded across ranks; every rank must name the same winners
        rng = np.random.default_rng(42)
        fitness = np.abs(rng.normal(size=(6, 8))) * 1e-5
        fitness[2, 3] = fitness[2, 6] = 1e-9     # an exact tie across ranks -> lowest guess id wins
        poses = rng.normal(size=(6, 8, 16)).astype(np.float32)
        s, c = pd.shard_range(8, rank, world)
        loc = fitness[:, s:s + c]
        arg = loc.argmin(axis=1)                  # first (lowest id) local minimum
        keys = pd.pack_key(loc[np.arange(6), arg], arg + s)
        k, p = pd.reduce_best(keys, poses[np.arange(6), arg + s])
        counts = pd.gather_frame_counts(pd.shard_range(1025, rank, world)[1])
        q.put((rank, k.tolist(), p.tolist(), counts))
    finally:
        dist.destroy_process_group()


def test_best_hypothesis_reduction_is_world_size_independent():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(42)
    fitness = np.abs(rng.normal(size=(6, 8))) * 1e-5
    fitness[2, 3] = fitness[2, 6] = 1e-9
    poses = rng.normal(size=(6, 8, 16)).astype(np.float32)
    want = fitness.argmin(axis=1)
    assert want[2] == 3
    for rank, k, p, counts in outs:
        got = [api.unpack_fitness_key(int(x))[1] for x in k]
        assert got == want.tolist()
        assert np.array_equal(np.asarray(p, np.float32), poses[np.arange(6), want])
        assert counts == [513, 512]
