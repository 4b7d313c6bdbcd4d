"""make_comm when ONE rank cannot set up its peer window: every rank must fall back to NCCL together (no hang), and
the collectives must still work.   torchrun --nproc-per-node 2 tests/multigpu_fallback_check.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.dist import PeerComm, TorchDistComm, make_comm


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ops = CudaOps(f"cuda:{local}")
    os.environ["ERA5SVD_COMM_TEST_FAIL_RANK"] = "1"
    comm = make_comm(ops)
    fell_back = isinstance(comm, TorchDistComm) and not isinstance(comm, PeerComm)
    t = torch.full((1000,), float(rank + 1), device="cuda", dtype=torch.float64)
    comm.allreduce_sum_(t)
    ok_sum = bool(torch.all(t == world * (world + 1) / 2))
    del os.environ["ERA5SVD_COMM_TEST_FAIL_RANK"]
    comm2 = make_comm(ops)                       # and without the simulated failure the peer path comes up
    peer_up = isinstance(comm2, PeerComm)
    u = torch.full((1000,), float(rank + 1), device="cuda", dtype=torch.float64)
    comm2.allreduce_sum_(u)
    ok_sum2 = bool(torch.all(u == world * (world + 1) / 2))
    flags = torch.tensor([fell_back, ok_sum, peer_up, ok_sum2], device="cuda", dtype=torch.float64)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "fell_back_on_every_rank": bool(flags[0]), "nccl_allreduce_ok": bool(flags[1]),
                          "peer_path_comes_up": bool(flags[2]), "peer_allreduce_ok": bool(flags[3])}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
