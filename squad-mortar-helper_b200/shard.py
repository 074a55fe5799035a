"""Frame-wise sharding across the GPUs of one box.

Frames are independent (the only cross-frame state in the reference is buffer reuse), so a batch is
split into static contiguous ranges, one per GPU, with NO data-path collective.  Inside one process the
C ABI does the split itself (smh_b200_multi_process_host_batch, one host thread + handle per GPU); under
torchrun (bench.py) every rank takes its range with the same arithmetic — this module is that arithmetic,
checked against smh_b200_shard_range in tests/test_sharding_cpu.py.
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
