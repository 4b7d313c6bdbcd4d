"""Parity metrics used by tests and bench (north_star tolerances).

  sigma : max_i |s_i - s_ref_i| / s_ref_i            <= 1e-6 (FP64) / 1e-4 (FP32 split)
  angle : per-vector principal angle, sign-free, asin(|| u - u_ref (u_ref^T u) ||) in
          float64 (accurate for small angles; acos(|u^T u_ref|) is not) <= 1e-5 rad (FP64)
  recon : || X - U S V ||_F / ||X||_F  within 1 % of the reference's
"""
from __future__ import annotations

import numpy as np


def sigma_rel_err(s, s_ref) -> float:
    s = np.asarray(s, dtype=np.float64); s_ref = np.asarray(s_ref, dtype=np.float64)
    return float(np.max(np.abs(s - s_ref) / np.abs(s_ref)))


def vector_angles(U, U_ref) -> np.ndarray:
    """Columns of U vs columns of U_ref; returns angle per column (radians)."""
    U = np.asarray(U, dtype=np.float64); U_ref = np.asarray(U_ref, dtype=np.float64)
    U = U / np.linalg.norm(U, axis=0); U_ref = U_ref / np.linalg.norm(U_ref, axis=0)
    c = np.sum(U * U_ref, axis=0)
    resid = U - U_ref * c
    return np.arcsin(np.minimum(1.0, np.linalg.norm(resid, axis=0)))


def subspace_angle(U, U_ref) -> float:
    """Largest principal angle between span(U) and span(U_ref)."""
    U = np.linalg.qr(np.asarray(U, dtype=np.float64))[0]
    U_ref = np.linalg.qr(np.asarray(U_ref, dtype=np.float64))[0]
    resid = U - U_ref @ (U_ref.T @ U)
    return float(np.arcsin(min(1.0, np.linalg.norm(resid, 2))))


def signs_agree(U, U_ref) -> bool:
    c = np.sum(np.asarray(U, np.float64) * np.asarray(U_ref, np.float64), axis=0)
    return bool(np.all(c > 0))


def recon_rel_err(X, U, s, Vt) -> float:
    X = np.asarray(X, dtype=np.float64)
    R = X - (np.asarray(U, np.float64) * np.asarray(s, np.float64)) @ np.asarray(Vt, np.float64)
    return float(np.linalg.norm(R) / np.linalg.norm(X))


def orthonormality(U) -> float:
    U = np.asarray(U, dtype=np.float64)
    return float(np.max(np.abs(U.T @ U - np.eye(U.shape[1]))))
