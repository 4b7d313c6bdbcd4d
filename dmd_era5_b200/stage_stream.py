"""Overlapped execution of the SVD stage over a SEQUENCE of host-resident ERA5 slices.

The reference handles one slice per ``era5_svd.main`` call (src/dmd_era5/era5_svd/era5_svd.py:336-453) and
is compute bound on the host.  On a B200 the stage itself (build + randomized SVD, ~17 ms for the
1 038 240 x 744 float32 slice) is several times shorter than moving that slice over PCIe (~55 ms), so a
production run over many slices (months of ERA5, variables, bootstrap resamples) is bound by the
host -> device copy unless the copies overlap the compute.  ``SvdStageStream`` does exactly that with three
CUDA streams and ``depth`` device input buffers:

    copy-in stream  : H2D of slice i+1 (pinned host memory -> device, native (T, S) layout)
    compute stream  : matrix build + SVD of slice i (the same calls as the one-shot path)
    copy-out stream : D2H of U, s, V of slice i-1 into pinned host buffers

Every slice still crosses PCIe in both directions; nothing is cached between slices.
"""
from __future__ import annotations

from typing import Callable, Iterable

import torch

from .device_ops import CudaOps
from .pipeline import build_matrix_device, svd_device


class SvdStageStream:
    def __init__(self, ops: CudaOps, T: int, S: int, *, n_components: int, svd_type: str = "randomized",
                 precision: str = "tf32x3", mean_center: bool = True, scale: bool = False, delay: int = 1,
                 seed: int | None = None, dtype: torch.dtype = torch.float32, depth: int = 2, comm=None,
                 row_offset: int = 0, m0_global: int | None = None):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one buffer being filled while another is being factorised)")
        self.ops, self.T, self.S, self.k = ops, int(T), int(S), int(n_components)
        self.svd_type, self.precision, self.mean_center, self.scale = svd_type, precision, mean_center, scale
        self.delay, self.seed, self.dtype, self.depth, self.comm = delay, seed, dtype, depth, comm
        self.row_offset, self.m0_global = row_offset, m0_global
        dev = ops.device
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        n = self.T - delay + 1
        kk = min(self.k, n)
        self.dev_in = [torch.empty((self.T, self.S), dtype=dtype, device=dev) for _ in range(depth)]
        self.host_out = [(torch.empty((self.S * delay, kk), dtype=dtype, pin_memory=True),
                          torch.empty((kk,), dtype=torch.float64, pin_memory=True),
                          torch.empty((kk, n), dtype=torch.float64, pin_memory=True)) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]       # H2D of the slot finished
        self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]      # compute on the slot finished
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]      # D2H of the slot's results finished
        self._keep = [None] * depth                                    # device results alive until their D2H is done
        self.h2d_bytes = self.T * self.S * self.dev_in[0].element_size()
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.host_out[0])

    def _compute(self, src: torch.Tensor):
        built = build_matrix_device(self.ops, [src], mean_center=self.mean_center, scale=self.scale)
        return svd_device(self.ops, built.X, svd_type=self.svd_type, n_components=self.k, delay=self.delay,
                          seed=self.seed, precision=self.precision, comm=self.comm, row_offset=self.row_offset,
                          m0_global=self.m0_global, centred=self.mean_center)

    def run(self, host_slices: Iterable[torch.Tensor],
            consume: Callable[[int, torch.Tensor, torch.Tensor, torch.Tensor], None] | None = None) -> int:
        """Factorise every (T, S) pinned host tensor of ``host_slices``.  ``consume(i, U, s, V)`` is called with
        the pinned host results of slice i once they have landed (the buffers are reused ``depth`` slices
        later, so copy what must be kept).  Returns the number of slices processed."""
        dev = self.ops.device
        it = iter(host_slices)
        pending: list[int] = []          # slice indices whose D2H has been enqueued but not consumed

        def enqueue_in(i, host):
            b = i % self.depth
            if i >= self.depth:
                self.s_in.wait_event(self.ev_cmp[b])          # the previous occupant of the slot has been read
            with torch.cuda.stream(self.s_in):
                self.dev_in[b].copy_(host, non_blocking=True)
                self.ev_in[b].record(self.s_in)

        def finish(i):
            b = i % self.depth
            self.ev_out[b].synchronize()
            self._keep[b] = None
            if consume is not None:
                consume(i, *self.host_out[b])

        nxt = next(it, None)
        i = 0
        if nxt is not None:
            enqueue_in(0, nxt)
        while nxt is not None:
            cur_i = i
            nxt = next(it, None)
            if nxt is not None:
                enqueue_in(cur_i + 1, nxt)                     # overlaps the compute of slice cur_i
            b = cur_i % self.depth
            if cur_i >= self.depth:
                finish(pending.pop(0))                         # frees host_out[b] and the kept device results
            self.s_cmp.wait_event(self.ev_in[b])
            with torch.cuda.device(dev), torch.cuda.stream(self.s_cmp):
                U, s, V = self._compute(self.dev_in[b])
                self.ev_cmp[b].record(self.s_cmp)
            self._keep[b] = (U, s, V)
            self.s_out.wait_event(self.ev_cmp[b])
            with torch.cuda.stream(self.s_out):
                hU, hs, hV = self.host_out[b]
                hU.copy_(U, non_blocking=True); hs.copy_(s, non_blocking=True); hV.copy_(V, non_blocking=True)
                self.ev_out[b].record(self.s_out)
            pending.append(cur_i)
            i += 1
        while pending:
            finish(pending.pop(0))
        return i
