"""Frame sharding for the multi-GPU path: contiguous frame ranges per rank, results gathered to rank 0.

Frames are independent in batch mode (ref: sample_compute.h:55-57 re-initialises the rings for every capture), so
there is no data-path collective; the only exchange is the gather of the small per-frame results.  Works with any
torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""


def frame_range(rank: int, world: int, n_frames: int):
    """Contiguous slice [lo, hi) of a global batch owned by `rank` (same rule as at_localize_host_sharded)."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def gather_to_rank0(local, n_frames: int, dist=None, dst: int = 0):
    """Gather per-rank result slices (first dimension = frames of that rank's range) into one tensor on `dst`.
    Slices may differ in length by one frame; they are padded to the longest for the collective."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [frame_range(r, world, n_frames) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
