"""``n_gpus`` > 1 behind the reference's own entry point: ``main(config)`` (src/dmd_era5/era5_svd/era5_svd.py:336-453)
row-shards the stage over N GPUs of this node without the caller changing anything but one opt-in config key.

One worker PROCESS per GPU (torch.multiprocessing spawn, NCCL process group on 127.0.0.1), the same per-rank driver as
``bench.py --gpus N``: every rank opens the slice file lazily, stages only its own rows (stage.shard_pieces: contiguous
base-row block, 128-row aligned), builds its shard of the snapshot matrix and runs the SVD with the NCCL communicator -
only the small n x l / l x l factors cross NVLink (dist.py).  U never leaves its rank on the device side; each rank writes
its rows to a scratch directory as .npy and the parent assembles the reference's layout (block-major over the delay
blocks).  The 236 GB configuration (BASELINE configs[2]) therefore runs through the drop-in API on >= 4 GPUs.
"""
from __future__ import annotations

import os
import shutil
import socket
import tempfile

import numpy as np


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, parsed_config: dict, outdir: str) -> None:
    import torch
    import torch.distributed as dist

    from .device_ops import CudaOps
    from .dist import make_comm
    from .stage import _device_arrays, _prepare, retrieve_era5_slice

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        ds, _ = retrieve_era5_slice(parsed_config, use_dvc=False)      # the parent has made sure the file is there
        if ds is None:
            raise FileNotFoundError(f"ERA5 slice {parsed_config['era5_slice_path']} not found by rank {rank}")
        ops = CudaOps(rank)
        comm = make_comm(ops)
        try:
            arr = _device_arrays(_prepare(ds, parsed_config), parsed_config, ops, comm=comm, rank=rank, world=world)
        finally:
            comm.close()              # barrier first: no rank frees its peer window while another may still read it
        np.save(os.path.join(outdir, f"U_{rank}.npy"), arr["U"])
        for key in ("X", "mean", "std"):
            if arr[key] is not None:
                np.save(os.path.join(outdir, f"{key}_{rank}.npy"), arr[key])
        np.save(os.path.join(outdir, f"rows_{rank}.npy"), np.array([*arr["rows"], arr["m0"], arr["S"]], dtype=np.int64))
        if rank == 0:
            np.save(os.path.join(outdir, "s.npy"), arr["s"])
            np.save(os.path.join(outdir, "V.npy"), arr["V"])
    finally:
        dist.destroy_process_group()


def assemble(outdir: str, world: int, d: int) -> dict:
    """Per-rank files -> the arrays of the single-device path (U rows block-major over the d delay blocks)."""
    meta = [np.load(os.path.join(outdir, f"rows_{r}.npy")) for r in range(world)]
    m0, S = int(meta[0][2]), int(meta[0][3])
    s, V = np.load(os.path.join(outdir, "s.npy")), np.load(os.path.join(outdir, "V.npy"))
    k = V.shape[0]
    U = None
    parts = {"X": None, "mean": None, "std": None}
    for r in range(world):
        r0, r1 = int(meta[r][0]), int(meta[r][1])
        Ul = np.load(os.path.join(outdir, f"U_{r}.npy"), mmap_mode="r")
        if U is None:
            U = np.empty((m0 * d, k), dtype=Ul.dtype)
        ml = r1 - r0
        for j in range(d):
            U[j * m0 + r0 : j * m0 + r1] = Ul[j * ml : (j + 1) * ml]
        for key in parts:
            f = os.path.join(outdir, f"{key}_{r}.npy")
            if os.path.exists(f):
                a = np.load(f, mmap_mode="r")
                if parts[key] is None:
                    parts[key] = np.empty((m0,) + a.shape[1:], dtype=a.dtype)
                parts[key][r0:r1] = a
    return {"U": U, "s": s, "V": V, "X": parts["X"], "mean": parts["mean"], "std": parts["std"], "m0": m0, "S": S}


def compute_multi(parsed_config: dict, n_gpus: int) -> dict:
    import torch
    import torch.multiprocessing as mp

    if not torch.cuda.is_available() or torch.cuda.device_count() < n_gpus:
        have = torch.cuda.device_count() if torch.cuda.is_available() else 0
        raise RuntimeError(f"n_gpus = {n_gpus} but only {have} CUDA device(s) are visible (no CPU fallback)")
    outdir = tempfile.mkdtemp(prefix="era5svd_multi_", dir=os.environ.get("ERA5SVD_SCRATCH"))
    try:
        mp.spawn(_worker, args=(n_gpus, _free_port(), dict(parsed_config), outdir), nprocs=n_gpus, join=True)
        return assemble(outdir, n_gpus, int(parsed_config["delay_embedding"]))
    finally:
        shutil.rmtree(outdir, ignore_errors=True)
