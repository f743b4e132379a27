"""Multi-GPU plumbing for the one exchange step of the path: summing reverse-mode parameter gradients.

Instances (pixels / parameter sets) are independent, so they are sharded across ranks in contiguous
blocks with no data-path collective; the value image stays sharded.  Only the
per-parameter gradient vector [P] (P = 6..2304 numbers) is combined, with ONE all-reduce per step.
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) carries it; the C ABI offers the
same operation without torch (tegb_nccl_comm_create / tegb_allreduce_sum).
"""
from __future__ import annotations

from typing import Tuple


def shard_rows(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rows owned by ``rank``: sizes differ by at most one, earlier ranks get the extra."""
    assert world_size >= 1 and 0 <= rank < world_size and n_rows >= 0
    base, extra = divmod(n_rows, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_instances(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    return shard_rows(n, world_size, rank)


def allreduce_gradients(grad_sum, group=None):
    """In-place sum of the local per-parameter gradient sums over all ranks (one collective)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad_sum, op=dist.ReduceOp.SUM, group=group)
    return grad_sum
