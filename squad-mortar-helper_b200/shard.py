"""Frame-wise sharding across the GPUs of one box.

Frames are independent (the only cross-frame state in the reference is buffer reuse), so a batch is
split into static contiguous ranges, one per rank, with NO data-path collective; ranks only meet at
a barrier and to combine their timings (max) and frame counts (sum).  `torch.distributed` is the
plumbing: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations


def shard_range(rank: int, world: int, total: int) -> tuple[int, int]:
    """Strong scaling: [first, last) of the `total` frames rank `rank` owns: g*B/G .. (g+1)*B/G."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    return rank * total // world, (rank + 1) * total // world


def weak_shard(rank: int, per_rank: int) -> tuple[int, int]:
    """Weak scaling (bench.py): every rank owns `per_rank` frames; global index = rank*per_rank + i."""
    return rank * per_rank, (rank + 1) * per_rank


def combine(frames_done: int, elapsed_ms: float, dist=None, device=None) -> tuple[int, float]:
    """(sum of frames over ranks, max of elapsed time over ranks).  `dist` is an initialised
    torch.distributed module or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return frames_done, elapsed_ms
    import torch
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    f = torch.tensor([frames_done], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return int(f.cpu()[0]), float(t.cpu()[0])


def frames_per_second(frames_done: int, elapsed_ms: float) -> float:
    return frames_done / (elapsed_ms * 1e-3)
