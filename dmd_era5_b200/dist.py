"""Row (space) sharding across GPUs: one process per GPU, ``torch.distributed`` for plumbing.

The snapshot matrix shards by rows - grid points / levels / variables are independent
 - so only time- or sketch-sized float64 factors ever cross NVLink:
  * all-reduce(sum) of Z = X^T Y        (n x l)     once per pass
  * all-reduce(sum) of G = Y^T Y        (l x l)     once
  * all-gather of the svd_flip candidates (k triples) once
Nothing space-sized is ever communicated; U stays sharded.

Two communicators over the same process group: ``TorchDistComm`` (NCCL / gloo through torch.distributed) and
``PeerComm`` (one node, CUDA devices): the small float64 collectives as OUR kernels over peer-mapped memory
(csrc/comm.cu) - the all-reduce of Z fused into the kernel that sums the projection's partial tiles - with
torch.distributed used only to exchange the IPC handles and for anything that does not fit a slot.
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

    def fuse_next_project(self, n: int, l: int) -> bool:
        return False

    def close(self) -> None:
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

    def fuse_next_project(self, n: int, l: int) -> bool:
        """True when the NEXT ops.project* call will leave the all-reduced result in Z (PeerComm only)."""
        return False

    def close(self) -> None:
        """Release communicator resources (PeerComm: barrier, then unmap the peers' windows and free this rank's - a rank
        must not free its window while a peer's last collective kernel may still be reading it)."""


class PeerComm(TorchDistComm):
    """Collectives as kernels of libera5svd.so over CUDA-IPC peer memory (NVLink / NVSwitch), one process per GPU of one
    node (era5svd_comm_*, include/era5svd.h).  float64 tensors that fit a slot take the peer path - one kernel on the
    compute stream, ranks summed in rank order (bit-identical replicas); everything else falls through to NCCL.
    Raises when peer access is not available: callers that want a fallback catch that and keep TorchDistComm."""

    def __init__(self, ops, group=None, slot_bytes: int = 4 << 20):
        import ctypes as C

        super().__init__(group)
        from ._cabi import check

        self.ops, self.lib = ops, ops.lib
        self._check = check
        self._comm = None
        hb = int(self.lib.era5svd_comm_handle_bytes())
        handle = (C.c_ubyte * hb)()
        comm = C.c_void_p()
        # Every rank walks through the SAME collectives whatever happens locally: a rank whose window cannot be created or
        # mapped reports that in an agreed all-reduce, and then all ranks raise together (no rank is left waiting inside
        # a collective the others never enter).
        with torch.cuda.device(ops.device):
            import os

            if os.environ.get("ERA5SVD_COMM_TEST_FAIL_RANK") == str(self.rank):      # test hook: this rank has no peer access
                rc, err = -2, "simulated failure (ERA5SVD_COMM_TEST_FAIL_RANK)"
            else:
                rc = self.lib.era5svd_comm_create(self.world, self.rank, int(slot_bytes), C.byref(comm), handle)
                err = "" if rc == 0 else self.lib.era5svd_last_error().decode(errors="replace")
            self._agree(rc == 0, "era5svd_comm_create", err, comm if rc == 0 else None)
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=ops.device)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            self._dist.all_gather(parts, mine, group=self.group)          # plumbing: the handle exchange
            allh = bytes(torch.cat(parts).cpu().tolist())
            buf = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
            rc = self.lib.era5svd_comm_connect(comm, buf)
            err = "" if rc == 0 else self.lib.era5svd_last_error().decode(errors="replace")
            self._agree(rc == 0, "era5svd_comm_connect", err, comm)
        self._comm = comm
        self.capacity = int(self.lib.era5svd_comm_capacity(self._comm))
        self._dist.barrier(group=self.group)

    def _agree(self, ok_here: bool, what: str, err: str, comm) -> None:
        flag = torch.tensor([1.0 if ok_here else 0.0], device=self.ops.device)
        self._dist.all_reduce(flag, op=self._dist.ReduceOp.MIN, group=self.group)
        if float(flag) < 1.0:
            if comm is not None:
                self.lib.era5svd_comm_destroy(comm)
            raise RuntimeError(f"PeerComm: {what} failed on " + (f"this rank: {err}" if not ok_here else "another rank"))

    def close(self) -> None:
        if getattr(self, "_comm", None):
            self._dist.barrier(group=self.group)
            self.lib.era5svd_comm_destroy(self._comm)
            self._comm = None

    def _fits(self, t: torch.Tensor, count: int) -> bool:
        return t.is_cuda and t.dtype == torch.float64 and 0 < count <= self.capacity

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if not (self._fits(t, t.numel()) and t.is_contiguous()):
            return super().allreduce_sum_(t)
        self._check(self.lib.era5svd_comm_allreduce_f64(self._comm, t.data_ptr(), t.numel(), self.ops._stream()),
                    "era5svd_comm_allreduce_f64")
        return t

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        t = t.contiguous()
        if not self._fits(t, t.numel()):
            return super().allgather(t)
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self._check(self.lib.era5svd_comm_allgather_f64(self._comm, t.data_ptr(), t.numel(), out.data_ptr(),
                                                        self.ops._stream()), "era5svd_comm_allgather_f64")
        return out

    def fuse_next_project(self, n: int, l: int) -> bool:
        if n * l > self.capacity:
            return False
        if self.ops.timer is not None:
            # a per-launch timer brackets the projection call (bench.py, every 4th step): keep the wait for the slowest
            # rank out of that bracket - the all-reduce then runs as its own kernel right after (same result, bit for bit)
            return False
        self._check(self.lib.era5svd_comm_fuse_next_project(self._comm, int(n), int(l)), "era5svd_comm_fuse_next_project")
        return True

    @property
    def fused_count(self) -> int:
        return int(self.lib.era5svd_comm_fused_count(self._comm))


def make_comm(ops, group=None, prefer_peer: bool = True):
    """Communicator for an initialised process group: PeerComm on CUDA devices of one node when peer memory can be mapped
    (every rank must succeed), else TorchDistComm.  ERA5SVD_COMM=nccl forces the torch.distributed path."""
    import os

    import torch.distributed as dist

    if not prefer_peer or os.environ.get("ERA5SVD_COMM", "peer") != "peer" or dist.get_backend(group) != "nccl":
        return TorchDistComm(group)
    try:
        return PeerComm(ops, group)       # raises on EVERY rank when any rank could not create / map its window
    except RuntimeError:                  # no peer access / IPC refused: NCCL carries the collectives
        return TorchDistComm(group)


def shard_rows(m0: int, world: int, rank: int, align: int = 128) -> tuple[int, int]:
    """Contiguous block of base rows [r0, r1) owned by ``rank``; shard edges are aligned to
    ``align`` rows (the GEMM row tile) except the last one."""
    if world <= 1:
        return 0, m0
    tiles = -(-m0 // align)
    base, extra = divmod(tiles, world)          # balanced: the first ``extra`` ranks own one tile more, so no rank is
    t0 = rank * base + min(rank, extra)         # left empty while tiles >= world (ceil(tiles / world) per rank left the
    t1 = t0 + base + (1 if rank < extra else 0)  # last ranks of a small matrix without rows: 21 tiles on 8 ranks)
    return min(m0, t0 * align), min(m0, t1 * align)


def shard_rows_weighted(m0: int, weights, rank: int, align: int = 128) -> tuple[int, int]:
    """Contiguous block of base rows owned by ``rank`` when the ranks' shares are proportional to ``weights`` (one
    positive number per rank, e.g. the rows per millisecond each GPU sustained in a calibration step: GPUs of one node
    differ by ~10 % under their power caps, and every collective waits for the slowest rank).  Shard edges are aligned to
    ``align`` rows except the last one; equal weights reproduce ``shard_rows`` up to rounding."""
    w = [max(float(x), 0.0) for x in weights]
    tot = sum(w)
    if tot <= 0 or len(w) <= 1:
        return shard_rows(m0, len(w), rank, align)
    tiles = -(-m0 // align)
    edges, acc = [0], 0.0
    for x in w:
        acc += x
        edges.append(max(edges[-1], min(tiles, int(round(acc / tot * tiles)))))
    edges[-1] = tiles
    return min(m0, edges[rank] * align), min(m0, edges[rank + 1] * align)
