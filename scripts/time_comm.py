"""Collective latencies: era5svd_comm_* (peer memory) against NCCL, and the projection with its all-reduce fused against
projection + NCCL.   torchrun --nproc-per-node N scripts/time_comm.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.dist import PeerComm, TorchDistComm


def timed(fn, reps=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 2)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ops = CudaOps(f"cuda:{local}")
    nccl, peer = TorchDistComm(), PeerComm(ops)
    out = {"world": world, "unit": "us per call, max over ranks"}
    for name, cnt in (("l x l (110 x 110)", 110 * 110), ("c2 Z (744 x 110)", 744 * 110), ("c3 Z (1460 x 110)", 1460 * 110),
                      ("flip candidates (3 x 100)", 300)):
        a = torch.randn(cnt, device="cuda", dtype=torch.float64)
        out[f"allreduce {name}"] = {"peer": timed(lambda: peer.allreduce_sum_(a)), "nccl": timed(lambda: nccl.allreduce_sum_(a))}
    b = torch.randn(300, device="cuda", dtype=torch.float64)
    out["allgather 300"] = {"peer": timed(lambda: peer.allgather(b)), "nccl": timed(lambda: nccl.allgather(b))}
    for (m, n, l) in ((262144, 744, 110), (262144, 1460, 110)):
        X = torch.randn((m, (n + 7) // 8 * 8), device="cuda")[:, :n]
        Y = torch.randn((m, ops.tf32_ldy(l)), device="cuda")[:, :l]
        Z = torch.zeros((n, l), device="cuda", dtype=torch.float64)

        def plain():
            ops.project_tf32x1(X, Y, Z)

        def with_nccl():
            ops.project_tf32x1(X, Y, Z); nccl.allreduce_sum_(Z)

        def with_peer_separate():
            ops.project_tf32x1(X, Y, Z); peer.allreduce_sum_(Z)

        def fused():
            assert peer.fuse_next_project(n, l)
            ops.project_tf32x1(X, Y, Z)

        out[f"project_x1 {m}x{n}x{l}"] = {"no collective": timed(plain, 50, 5), "+ nccl all-reduce": timed(with_nccl, 50, 5),
                                          "+ peer all-reduce kernel": timed(with_peer_separate, 50, 5), "fused": timed(fused, 50, 5)}
    peer.close()
    if rank == 0:
        print(json.dumps(out, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
