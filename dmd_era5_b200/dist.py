"""Row (space) sharding across GPUs: one process per GPU, ``torch.distributed`` for plumbing.

The snapshot matrix shards by rows - grid points / levels / variables are independent
 - so only time- or sketch-sized float64 factors ever cross NVLink:
  * all-reduce(sum) of Z = X^T Y        (n x l)     once per pass
  * all-reduce(sum) of G = Y^T Y        (l x l)     once
  * all-gather of the svd_flip candidates (k triples) once
Nothing space-sized is ever communicated; U stays sharded.
"""
from __future__ import annotations

import torch


class LocalComm:
    """Single-process communicator (world size 1)."""

    rank = 0
    world = 1

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        return t.unsqueeze(0)

    def barrier(self) -> None:
        pass


class TorchDistComm:
    """Communicator over an initialised ``torch.distributed`` process group (NCCL on GPUs,
    gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if t.is_contiguous():
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
            return t
        tmp = t.contiguous()
        self._dist.all_reduce(tmp, op=self._dist.ReduceOp.SUM, group=self.group)
        t.copy_(tmp)
        return t

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        t = t.contiguous()
        parts = [torch.empty_like(t) for _ in range(self.world)]
        self._dist.all_gather(parts, t, group=self.group)
        return torch.stack(parts)

    def barrier(self) -> None:
        self._dist.barrier(group=self.group)


def shard_rows(m0: int, world: int, rank: int, align: int = 128) -> tuple[int, int]:
    """Contiguous block of base rows [r0, r1) owned by ``rank``; shard edges are aligned to
    ``align`` rows (the GEMM row tile) except the last one."""
    if world <= 1:
        return 0, m0
    tiles = -(-m0 // align)
    per = -(-tiles // world)
    r0 = min(m0, rank * per * align)
    r1 = min(m0, (rank + 1) * per * align)
    return r0, r1
