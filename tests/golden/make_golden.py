"""Generate the golden fixtures in this directory.  Run ONCE in the build container
(where /root/reference exists); the GPU box only reads the committed .npz files.

    python tests/golden/make_golden.py

Sources of truth:
  * delay_embedding.npz - produced by the reference's OWN ``_apply_delay_embedding_np``
    (src/dmd_era5/slice_tools/slice_tools.py:182-211), whose source is extracted with
    ``ast`` and executed here (the package itself cannot be imported: xarray, dvc and
    pyprojroot are absent from this image).  Also holds the 4 known-answer cases of
    tests/test_02_slice_tools.py:215-231.
  * svd_standard_c1.npz / svd_randomized_*.npz - outputs of the exact library calls the
    reference makes (era5_svd.py:251, :258) on seeded inputs that tests regenerate.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference/src/dmd_era5/slice_tools/slice_tools.py"


def load_reference_delay_fn():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_apply_delay_embedding_np")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np}
    from numpy.lib.stride_tricks import sliding_window_view
    ns["sliding_window_view"] = sliding_window_view
    exec(compile(mod, REF, "exec"), ns)
    return ns["_apply_delay_embedding_np"]


def main():
    from oracle.svd_ref import randomized_svd_ref, standard_svd_ref
    from oracle.slice_tools_np import build_matrix_np
    from oracle.synthetic_np import lowrank_field_np, mock_era5_np

    # ---- delay embedding, from the reference's own function --------------------
    ref_delay = load_reference_delay_fn()
    out = {}
    kat = [
        (np.array([[0, 1, 2, 3, 4]]), 1), (np.array([[0, 1, 2, 3, 4]]), 2),
        (np.array([[0, 1, 2, 3, 4]]), 3), (np.array([[0, 1, 2], [3, 4, 5]]), 2),
    ]
    for i, (X, d) in enumerate(kat):
        out[f"kat{i}_X"] = X; out[f"kat{i}_d"] = d; out[f"kat{i}_out"] = ref_delay(X, d)
    rng = np.random.RandomState(11)
    for i, (m, T, d) in enumerate([(7, 9, 1), (7, 9, 4), (33, 25, 2), (5, 6, 6)]):
        X = rng.standard_normal((m, T))
        out[f"rnd{i}_X"] = X; out[f"rnd{i}_d"] = d; out[f"rnd{i}_out"] = ref_delay(X, d)
    np.savez_compressed(os.path.join(HERE, "delay_embedding.npz"), **out)

    # ---- config 1: mock ERA5, standard SVD (np.linalg.svd, era5_svd.py:251) -----
    ds = mock_era5_np(25, ["temperature", "u_component_of_wind"], [1000], seed=0)
    X, _, _ = build_matrix_np([ds["vars"][v] for v in ds["vars"]], True, False, 2)
    U, s, V = standard_svd_ref(X, 6)
    np.savez_compressed(os.path.join(HERE, "svd_standard_c1.npz"), U=U, s=s, V=V,
                        meta=np.array([25, 2, 1, 2, 6, 0]))  # T, nvars, nlev, d, k, seed

    # ---- config 1 input, randomized (sklearn, era5_svd.py:258), seed 5 ----------
    U, s, V = randomized_svd_ref(X, 6, 5)
    np.savez_compressed(os.path.join(HERE, "svd_randomized_c1.npz"), U=U, s=s, V=V,
                        meta=np.array([25, 2, 1, 2, 6, 0, 5]))

    # ---- low-rank separated spectrum, randomized, f64 and f32 inputs ------------
    X = lowrank_field_np(2048, 96, r=40, rho=0.8, seed=2)
    U, s, V = randomized_svd_ref(X, 12, 9)
    np.savez_compressed(os.path.join(HERE, "svd_randomized_lowrank_f64.npz"), U=U, s=s, V=V,
                        meta=np.array([2048, 96, 40, 12, 2, 9]))
    U, s, V = randomized_svd_ref(X.astype(np.float32), 12, 9)
    np.savez_compressed(os.path.join(HERE, "svd_randomized_lowrank_f32.npz"), U=U, s=s, V=V,
                        meta=np.array([2048, 96, 40, 12, 2, 9]))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
