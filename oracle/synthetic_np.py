"""Seeded CPU generators for test inputs (test infrastructure only).

mock_era5_np restates the reference's create_mock_data generator
(src/dmd_era5/create_mock_data/create_mock_data.py:65-71, 135-149) with a seed:
5-degree grid (36 x 72), hourly inclusive range, float64,
temperature = (rand*30 + 250 - (1000-level)/100) * cos(lat), winds = rand*20 - 10.
"""
from __future__ import annotations

import numpy as np


def mock_era5_np(n_times: int, variables: list[str], levels: list[int], seed: int = 0):
    rng = np.random.RandomState(seed)
    lats = np.arange(90, -90, -5.0)
    lons = np.arange(-180, 180, 5.0)
    shape = (n_times, len(levels), len(lats), len(lons))
    out = {}
    for var in variables:
        if var == "temperature":
            data = rng.rand(*shape) * 30 + 250
            for i, level in enumerate(levels):
                data[:, i, :, :] -= (1000 - level) / 100
            data = data * np.cos(np.radians(lats))[np.newaxis, np.newaxis, :, np.newaxis]
        elif "wind" in var:
            data = rng.rand(*shape) * 20 - 10
        else:
            data = rng.rand(*shape) * 100
        out[var] = data
    return {"vars": out, "level": np.asarray(levels), "latitude": lats, "longitude": lons,
            "time": np.datetime64("2019-01-01T00", "ns") + np.arange(n_times) * np.timedelta64(1, "h")}


def lowrank_field_np(m: int, n: int, r: int = 160, rho: float = 0.93, sigma0: float = 100.0,
                     noise: float = 1e-5, seed: int = 0, dtype=np.float64) -> np.ndarray:
    """ERA5-shaped (space x time) matrix with a geometrically separated spectrum
    sigma_i = sigma0 * rho**i (>= 7 % gaps so per-vector parity is well posed,
    SURVEY 8d) plus a white-noise floor well below sigma_r."""
    rng = np.random.RandomState(seed)
    A = np.linalg.qr(rng.standard_normal((m, r)))[0]
    B = np.linalg.qr(rng.standard_normal((n, r)))[0]
    s = sigma0 * rho ** np.arange(r)
    X = (A * s) @ B.T + noise * sigma0 * rng.standard_normal((m, n)) / np.sqrt(m)
    return X.astype(dtype)
