"""BASELINE configs[3] (c4): 1 deg x 3 variables x 37 levels (7 232 760 rows) x 8760 hourly snapshots, float32 (253 GB), full
standard SVD by the Gram route, row-sharded over the ranks of one node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_c4.py [k] [rows_total]

Per-phase device times (CUDA events on the launching stream, max over ranks): Gram (tensor-core column blocks), the n x n
float64 all-reduce (614 MB), the replicated eigensolve (tridiagonalisation + top-k), the float32 refinement passes and
U.  Size-independent checks at full size: orthonormal U / V, X^T U = V^T S on rank 0's shard, descending sigma, and the
generator's spectrum."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from dmd_era5_b200 import standard as std_mod
from dmd_era5_b200.device_ops import CudaOps, KernelTimer
from dmd_era5_b200.dist import LocalComm, make_comm, shard_rows
from dmd_era5_b200.pipeline import build_matrix_device, svd_device
from dmd_era5_b200.synthetic import synthetic_field


class TimedComm:
    """Communicator wrapper that brackets every collective with CUDA events."""

    def __init__(self, inner):
        self.inner, self.rank, self.world, self.events = inner, inner.rank, inner.world, []

    def allreduce_sum_(self, t):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = self.inner.allreduce_sum_(t); e1.record()
        self.events.append((t.numel() * t.element_size(), e0, e1))
        return out

    def allgather(self, t):
        return self.inner.allgather(t)

    def fuse_next_project(self, n, l):
        return self.inner.fuse_next_project(n, l)

    def close(self):
        self.inner.close()

    def barrier(self):
        self.inner.barrier()


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 3 * 37 * 181 * 360
    T = 8760
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ops = CudaOps(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = TimedComm(make_comm(ops))
    else:
        comm = TimedComm(LocalComm())
    r0, r1 = shard_rows(M, world, rank)
    field = synthetic_field(T, r1 - r0, device=dev, seed=40 + rank, rank=200, rho=0.96, chunk=1 << 17, time_seed=40,
                            total_points=M)
    built = build_matrix_device(ops, [field], mean_center=True, scale=False)
    del field
    torch.cuda.empty_cache()
    X = built.X
    out = {}
    for rep in range(2):
        comm.events.clear()
        ops.timer = KernelTimer()
        torch.cuda.synchronize(); comm.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        stats = {}
        U, s, V = svd_device(ops, X, svd_type="standard", n_components=k, precision="auto", comm=comm, row_offset=r0,
                             m0_global=M, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        summ = ops.timer.summary(); ops.timer = None
        ms = torch.tensor([e0.elapsed_time(e1), sum(a.elapsed_time(b) for _, a, b in comm.events)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out = {"rep": rep, "world": world, "rows_total": M, "rows_this_rank": r1 - r0, "snapshots": T, "k": k,
               "matrix_GB": M * T * 4 / 1e9, "ms_total_max_over_ranks": float(ms[0]), "wall_s_rank0": wall,
               "GBps": M * T * 4 / 1e9 / (float(ms[0]) / 1e3),
               "allreduce_ms_max_over_ranks": float(ms[1]),
               "allreduce_bytes": [b for b, _, _ in comm.events],
               "kernels_ms_rank0": {kk: round(v["ms"], 2) for kk, v in summ.items()},
               "kernel_calls_rank0": {kk: v["calls"] for kk, v in summ.items()},
               "sigma_first3": s[:3].tolist(), "refine_iters": std_mod.REFINE_ITERS, "eig_route": stats.get("eig_route")}
    # invariants on this rank's shard
    Ud = U.double()
    G = Ud.t() @ Ud
    if world > 1:
        dist.all_reduce(G)
    eye = torch.eye(U.shape[1], device=dev, dtype=torch.float64)
    resid = torch.zeros((T, U.shape[1]), device=dev, dtype=torch.float64)
    for a in range(0, X.shape[0], 1 << 16):
        resid += X[a : a + (1 << 16)].double().t() @ Ud[a : a + (1 << 16)]
    if world > 1:
        dist.all_reduce(resid)
    out["UtU_minus_I_max"] = float((G - eye).abs().max())
    out["VVt_minus_I_max"] = float((V.double() @ V.double().t() - eye).abs().max())
    out["XtU_minus_VtS_rel"] = float((resid - V.double().t() * s.double()).norm() / s.double().norm())
    out["sigma_descending"] = bool(torch.all(s[:-1] >= s[1:]))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
