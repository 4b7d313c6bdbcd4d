"""Synthetic ERA5-shaped fields generated ON THE DEVICE (bench / large-size tests only).

Value model: field[t, s] = mu(s) + sum_{i<r} sigma_i a_i(s) b_i(t) + noise, with a
geometrically separated spectrum sigma_i = sigma0 * rho**i so that per-vector parity is well posed,
a_i smooth random spatial patterns, b_i random temporal patterns, a small white-noise floor, and a
temperature-like mean field.  torch's device RNG is used only to fabricate INPUT data; it is not
part of the SVD path.  Layout is the native ERA5 one: (T, S) time-major, float32 by default.
"""
from __future__ import annotations

import math

import torch


def synthetic_field(T: int, S: int, *, device, dtype=torch.float32, rank: int = 160, rho: float = 0.93,
                    sigma0: float = 100.0, noise: float = 1e-5, mean_level: float = 250.0, seed: int = 0,
                    chunk: int = 1 << 20, time_seed: int | None = None, total_points: int | None = None) -> torch.Tensor:
    """(T, S) field whose time-centred version has singular values ~ sigma0 * rho**i.

    Row-sharded runs: every rank passes the SAME ``time_seed`` (shared temporal patterns) and its own ``seed`` (its own
    spatial patterns) plus ``total_points`` = the global number of points, so that the shards together form ONE field with
    the designed spectrum; with per-rank temporal patterns the stacked matrix would instead carry world-size interleaved
    copies of the spectrum (sigma_i clustered in groups)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    gt = g
    if time_seed is not None:
        gt = torch.Generator(device=device)
        gt.manual_seed(time_seed)
    # temporal patterns: orthonormal columns (T x rank), exactly zero time mean
    Bt = torch.randn((T, rank), generator=gt, device=device, dtype=torch.float64)
    Bt -= Bt.mean(dim=0, keepdim=True)
    Bt, _ = torch.linalg.qr(Bt)
    s = sigma0 * rho ** torch.arange(rank, device=device, dtype=torch.float64)
    BtS = (Bt * s).to(torch.float32 if dtype == torch.float32 else torch.float64)      # (T, rank)
    out = torch.empty((T, S), device=device, dtype=dtype)
    inv_sqrt_S = 1.0 / math.sqrt(total_points or S)
    for c0 in range(0, S, chunk):
        c1 = min(S, c0 + chunk)
        # spatial patterns: i.i.d. N(0, 1/S) columns are orthonormal up to O(sqrt(rank/S))
        A = torch.randn((rank, c1 - c0), generator=g, device=device, dtype=BtS.dtype) * inv_sqrt_S
        blk = BtS @ A
        blk += noise * sigma0 * inv_sqrt_S * torch.randn((T, c1 - c0), generator=g, device=device, dtype=BtS.dtype)
        mu = mean_level + 30.0 * torch.cos(torch.linspace(-1.5, 1.5, c1 - c0, device=device, dtype=BtS.dtype))
        blk += mu
        out[:, c0:c1] = blk.to(dtype)
    return out
