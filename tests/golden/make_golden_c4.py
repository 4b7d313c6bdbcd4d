"""Golden fixture for the config-4 SHAPE (n = 8760 hourly snapshots, standard SVD) at reduced rows.

    python tests/golden/make_golden_c4.py        (about two minutes on 8 cores; run once in the build container)

Source of truth: the reference's own call for svd_type = "standard"
(src/dmd_era5/era5_svd/era5_svd.py:251-254): ``np.linalg.svd(X, full_matrices=False)`` then truncation, on a seeded
float64 matrix the test regenerates (oracle.synthetic_np.lowrank_field_np: 9600 x 8760, sigma_i = 100 * 0.9**i, r = 300).
Kept small: all K_SIGMA singular values, the K_VEC leading singular-vector pairs.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

M, N, R, RHO, SEED = 9600, 8760, 300, 0.9, 7
K_SIGMA, K_VEC = 48, 12


def main():
    from oracle.svd_ref import standard_svd_ref
    from oracle.synthetic_np import lowrank_field_np

    X = lowrank_field_np(M, N, r=R, rho=RHO, seed=SEED)
    t0 = time.perf_counter()
    U, s, V = standard_svd_ref(X, K_SIGMA)
    print(f"np.linalg.svd {M} x {N}: {time.perf_counter() - t0:.1f} s")
    np.savez_compressed(os.path.join(HERE, "svd_standard_n8760.npz"), s=s, U=U[:, :K_VEC], V=V[:K_VEC],
                        meta=np.array([M, N, R, SEED, K_SIGMA, K_VEC]), rho=np.array(RHO))
    print(os.path.getsize(os.path.join(HERE, "svd_standard_n8760.npz")), "bytes")


if __name__ == "__main__":
    main()
