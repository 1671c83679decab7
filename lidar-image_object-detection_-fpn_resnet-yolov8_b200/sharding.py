"""Frame sharding across the GPUs of one box.

Every sweep is independent, so the path shards by frame index with NO collective on the data
path: rank r of G processes frames r, r+G, r+2G, ... (or a contiguous block).  The only optional
exchange is a gather of the [n_local, K, 10] detections (2 KB per frame) to every rank, which goes
through torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world_size, mode="block"):
    """Indices of the frames rank `rank` owns.  block: contiguous, sizes differ by at most one;
    cyclic: i % world_size == rank."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    if mode == "cyclic":
        return range(rank, n_frames, world_size)
    if mode != "block":
        raise ValueError(mode)
    base, rem = divmod(n_frames, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_detections(local_det, n_frames, mode="block", group=None):
    """all-gather of per-rank detections [n_local, K, 10] into [n_frames, K, 10] in global frame order
    (ragged shard sizes are padded to the largest shard for the collective)."""
    if not dist.is_initialized():
        return local_det
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    K, D = local_det.shape[1], local_det.shape[2]
    n_max = max(len(shard_range(n_frames, r, world, mode)) for r in range(world))
    padded = local_det.new_zeros((n_max, K, D))
    padded[: local_det.shape[0]] = local_det
    buf = local_det.new_empty((world * n_max, K, D))
    dist.all_gather_into_tensor(buf, padded, group=group) if local_det.is_cuda else \
        dist.all_gather(list(buf.view(world, n_max, K, D).unbind(0)), padded, group=group)
    out = local_det.new_empty((n_frames, K, D))
    buf = buf.view(world, n_max, K, D)
    for r in range(world):
        idx = shard_range(n_frames, r, world, mode)
        if len(idx):
            out[torch.as_tensor(list(idx), device=out.device)] = buf[r, : len(idx)]
    return out
