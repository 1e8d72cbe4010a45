"""Data-parallel sharding of a triplet batch over the GPUs of one box.

Every (probe, mate, non-mate) triplet is independent (the reference farms them out per process,
eval/generate_inpaintinggame_wb_saliency_maps_multigpu.py:121-131,191-231), so rank r sweeps a
contiguous slice of the batch with replicated weights and the only collective on the data path is
the gather of the finished [n,112,112] fp32 maps (50 KB each) to rank 0.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n items for `rank`; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_maps(local_maps, n_total, dst=0, group=None):
    """Gather per-rank [n_r, H, W] maps to `dst` in global triplet order.  Ragged shards are padded to the
    largest shard for the collective and trimmed afterwards.  Returns the [n_total, H, W] tensor on dst,
    None elsewhere.  Works with NCCL (device tensors) and gloo (CPU tensors, used by the tests)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_maps
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = local_maps
    if local_maps.shape[0] < nmax:
        pad = torch.zeros((nmax,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
        pad[:local_maps.shape[0]] = local_maps
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)])
