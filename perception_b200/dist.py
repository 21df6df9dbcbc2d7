"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed for the
rendezvous. Frames are independent, so ranks take contiguous blocks of frames and no data-path collective
exists. The one real exchange is the low-latency mode that splits the ICP initial-pose hypotheses of a
cluster across ranks: each rank reports (fitness, guess id, pose) for its best local hypothesis, a single
all-gather of 80-byte records follows, and every rank picks the global winner as the exact lexicographic
minimum of (fitness, guess id) (reduce_hypotheses -> cuboid_reduce_guess_records) — so the answer does not depend
on the GPU count. pack_key / reduce_best are the older truncated-key variant (one MIN all-reduce), kept for callers
that can live with fitness values agreeing to 2^-36 being ordered by guess id."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [start, start+count) of n_items for `rank` (blocks differ by at most one item)."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def pack_key(fitness, guess_id):
    """Same packing as cuboid_pack_fitness_key (include/cuboid_cuda.h), vectorised; fits a signed int64."""
    f = np.asarray(fitness, dtype=np.float64).copy()
    f[~(f >= 0)] = np.finfo(np.float64).max
    b = f.view(np.uint64) & np.uint64(0xFFFFFFFFFFFF0000)
    return (b | (np.asarray(guess_id, dtype=np.uint64) & np.uint64(0xFFFF))).astype(np.int64)


def reduce_best(local_keys, local_poses, group=None):
    """local_keys int64 [n]; local_poses float32 [n,16]  ->  (winner keys [n], winner poses [n,16]) on every rank.

    One all_gather of (key, pose) records; the arg-min over ranks is taken locally from the keys."""
    import torch
    import torch.distributed as dist
    keys = torch.as_tensor(np.asarray(local_keys, dtype=np.int64))
    poses = torch.as_tensor(np.asarray(local_poses, dtype=np.float32)).reshape(len(keys), 16)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return keys.numpy(), poses.numpy()
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    rec = torch.cat([keys.view(-1, 1), poses.to(torch.float64).view(torch.int64)], dim=1).to(dev)   # [n, 17] int64 payload
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    allrec = torch.stack(out).cpu()                # [world, n, 17]
    win = allrec[:, :, 0].argmin(dim=0)            # keys are unique per guess id, so argmin is unambiguous
    idx = torch.arange(allrec.shape[1])
    best = allrec[win, idx]
    return best[:, 0].numpy(), best[:, 1:].contiguous().view(torch.float64).to(torch.float32).numpy()


def reduce_hypotheses(results, gate, group=None):
    """Hypothesis-sharded mode (second axis): every rank ran the SAME frames with its own slice of the initial-pose
    hypotheses (api.CuboidCuda.set_guesses(slice) + set_guess_offset(first id)). `results` = this rank's FrameResult array.
    One all_gather of the 80-byte cuboid_guess_record of every (frame, cluster), then the exact (fitness, guess id) arg-min of
    cuboid_reduce_guess_records, in place: afterwards every rank holds the results a single GPU running all hypotheses returns
    (corr_hash, a per-rank parity tap, is cleared). Returns `results`."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import api
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return results
    world = dist.get_world_size(group)
    n = len(results)
    recs = (api.GuessRecord * (n * api.MAX_CLUSTERS))()
    for f in range(n):
        for c in range(min(results[f].n_clusters, api.MAX_CLUSTERS)):
            api.load().cuboid_guess_record_from_result(C.byref(results[f].cluster[c]), C.byref(recs[f * api.MAX_CLUSTERS + c]))
    local = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    local = local.to(dev)
    out = torch.empty((world, local.numel()), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, local, group=group) if dev.type == "cuda" else dist.all_gather(list(out.unbind(0)), local, group=group)
    allb = out.cpu().numpy()
    rs = C.sizeof(api.GuessRecord)
    per = (api.GuessRecord * world)()
    for f in range(n):
        for c in range(min(results[f].n_clusters, api.MAX_CLUSTERS)):
            o = (f * api.MAX_CLUSTERS + c) * rs
            for r in range(world):
                C.memmove(C.byref(per[r]), allb[r, o:o + rs].tobytes(), rs)
            st = api.load().cuboid_reduce_guess_records(per, world, float(gate), C.byref(results[f].cluster[c]))
            if st != api.OK:
                raise api.CuboidError(st, "cuboid_reduce_guess_records")
    return results


def gather_frame_counts(local_count, group=None):
    """All ranks learn how many frames every rank processed (used to assemble whole-job throughput)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(local_count)]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [int(x.item()) for x in out]
