"""Environment-parallel (data-parallel) plumbing for GCBF.train_step -- SURVEY section 8e.

Graphs of a batch are independent (block-diagonal collation), so the B graphs are partitioned contiguously over the
ranks; weights, Adam state and the spectral-norm u/v buffers are replicated and stay bit-identical because every rank
applies the same all-reduced gradient.  Per step there are exactly three exchanges, all via torch.distributed (NCCL on
GPUs over NVLink/NVSwitch, gloo in the CPU tests):
  1. all-reduce of the 16 loss partial sums (fp64)   -> global counts for the masked means (gcbf.py:172,184,208,212)
  2. all-gather of h_dot (M floats) + all-reduce of one int64 pair count   -> the M x M `acc/derivative` (gcbf.py:209)
  3. ONE all-reduce (sum) of the flat fp32 gradient bucket of both nets (24.46 M floats)   -> clip + Adam (gcbf.py:220-226)
"""
from typing import Optional, Tuple

import torch


def shard_range(num_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of `num_items` graphs: rank r owns [start, stop)."""
    base, rem = divmod(num_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class Reducer:
    """Thin wrapper over a process group; a no-op for a single process."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0

    def sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def max_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def gather_cat(self, t: torch.Tensor) -> torch.Tensor:
        """Concatenation of equally sized per-rank vectors, in rank order."""
        if self.world == 1:
            return t
        out = torch.empty(self.world * t.numel(), device=t.device, dtype=t.dtype)
        self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out
