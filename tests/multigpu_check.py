"""Row-sharded randomized SVD over NCCL vs the single-process oracle.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_check.py

Every rank builds the same host matrix, keeps its row shard (delay-embedded, d = 2), runs the device
driver with the NCCL communicator, and rank 0 compares the gathered U / s / V with sklearn's
randomized_svd (float64) on the full matrix."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from dmd_era5_b200.device_ops import CudaOps
from dmd_era5_b200.dist import PeerComm, TorchDistComm, shard_rows
from dmd_era5_b200.pipeline import svd_device
from oracle.compare import sigma_rel_err, signs_agree, vector_angles
from oracle.slice_tools_np import delay_embed_np
from oracle.svd_ref import randomized_svd_ref
from oracle.synthetic_np import lowrank_field_np


def check_peer_collectives(peer, nccl, rank, world):
    """era5svd_comm_allreduce_f64 / allgather_f64 against NCCL on random data: many sizes back to back (slot parity,
    epoch flags), one rank delayed (the others must wait in the kernel), replicas bit-identical."""
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    worst, equal, gathered_ok = 0.0, True, True
    sizes = [1, 7, 110 * 110, 744 * 110, 1460 * 110, 1023, 1024, 1025, 3 * 24 * 2] * 6
    for it, cnt in enumerate(sizes):
        a = torch.randn(cnt, generator=g, device="cuda", dtype=torch.float64)
        ref = a.clone()
        if it % 5 == rank % 5:
            torch.cuda._sleep(2_000_000)                      # ~1 ms: this rank arrives late
        peer.allreduce_sum_(a)
        nccl.allreduce_sum_(ref)
        worst = max(worst, float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-300)))
        allv = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(allv, a)
        equal = equal and all(torch.equal(allv[0], t) for t in allv)
        b = torch.randn(cnt, generator=g, device="cuda", dtype=torch.float64)
        gathered_ok = gathered_ok and torch.equal(peer.allgather(b), nccl.allgather(b))
    big = torch.ones(peer.capacity + 5, device="cuda", dtype=torch.float64)          # does not fit a slot: NCCL carries it
    peer.allreduce_sum_(big)
    res = {"allreduce_max_rel_diff_vs_nccl": worst, "replicas_bit_identical": equal, "allgather_equal": gathered_ok,
           "oversize_falls_through": bool(torch.all(big == world)), "calls": len(sizes)}
    res["pass"] = bool(worst < 1e-14 and equal and gathered_ok and res["oversize_falls_through"])
    return res


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ops = CudaOps(f"cuda:{local}")
    nccl = TorchDistComm()
    peer = PeerComm(ops)              # raises if peer memory cannot be mapped: this check is about that path
    out = {}
    out["peer_collectives"] = check_peer_collectives(peer, nccl, rank, world)
    m0, T, k, d = 8192 * world + 77, 300, 24, 2
    X = lowrank_field_np(m0, T, r=80, rho=0.9, seed=11, dtype=np.float32)
    r0, r1 = shard_rows(m0, world, rank)
    results = {}
    for comm_kind, comm, precision, tol in (("nccl", nccl, "native", 1e-4), ("nccl", nccl, "tf32x3", 1e-4),
                                            ("nccl", nccl, "tf32mix", 1e-4), ("peer", peer, "native", 1e-4),
                                            ("peer", peer, "tf32mix", 1e-4)):
        key = precision if comm_kind == "nccl" else f"peer_{precision}"
        fused0 = peer.fused_count
        Xd = torch.from_numpy(X[r0:r1].copy()).cuda()
        U, s, V = svd_device(ops, Xd, svd_type="randomized", n_components=k, delay=d, seed=5, precision=precision,
                             comm=comm, row_offset=r0, m0_global=m0)
        # gather the (unequal) shards of U on rank 0: pad to the largest shard
        ml = torch.tensor([r1 - r0], device="cuda")
        sizes = [torch.zeros_like(ml) for _ in range(world)]
        dist.all_gather(sizes, ml)
        mx = int(max(x.item() for x in sizes))
        pad = torch.zeros((d, mx, k), device="cuda", dtype=U.dtype)
        pad[:, : r1 - r0] = U.reshape(d, r1 - r0, k)
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        # the replicated factors must be BIT-identical on every rank (sums formed in the same order everywhere)
        sv = torch.cat([s.reshape(-1), V.reshape(-1)]).contiguous()
        allsv = [torch.empty_like(sv) for _ in range(world)]
        dist.all_gather(allsv, sv)
        replicas_equal = all(torch.equal(allsv[0], t) for t in allsv)
        results[key] = (s.clone(), V.clone())
        if rank == 0:
            Ufull = np.zeros((m0 * d, k))
            for rk, (p, sz) in enumerate(zip(parts, sizes)):
                a0, a1 = shard_rows(m0, world, rk)
                for j in range(d):
                    Ufull[j * m0 + a0 : j * m0 + a1] = p[j, : int(sz.item())].double().cpu().numpy()
            U0, s0, V0 = randomized_svd_ref(delay_embed_np(X.astype(np.float64), d), k, 5)
            out[key] = {"sigma_rel_err": sigma_rel_err(s.cpu().numpy(), s0),
                        "angle_U_max": float(vector_angles(Ufull, U0).max()),
                        "angle_V_max": float(vector_angles(V.cpu().numpy().T, V0.T).max()),
                        "signs_agree": signs_agree(Ufull, U0), "replicas_bit_identical": replicas_equal}
            out[key]["pass"] = bool(out[key]["sigma_rel_err"] < tol and out[key]["signs_agree"]
                                    and out[key]["angle_U_max"] < 2e-2 and replicas_equal)
            if comm_kind == "peer":
                # (2 q + 2) / 2 fused projections (one per tall pass pair) + the Gram matrix, per call
                out[key]["fused_allreduces"] = peer.fused_count - fused0
                sn, Vn = results[precision]
                out[key]["sigma_vs_nccl_rel"] = float(((s - sn).abs() / sn).max())
                out[key]["pass"] = bool(out[key]["pass"] and out[key]["fused_allreduces"] > 0
                                        and out[key]["sigma_vs_nccl_rel"] < 1e-9)
    # standard SVD (Gram route: the n x n Gram matrix is the only collective), float64, against np.linalg.svd
    from oracle.svd_ref import standard_svd_ref

    X64 = lowrank_field_np(4096 * world + 33, 200, r=60, rho=0.9, seed=12)
    a0, a1 = shard_rows(X64.shape[0], world, rank)
    comm = nccl
    U, s, V = svd_device(ops, torch.from_numpy(X64[a0:a1].copy()).cuda(), svd_type="standard", n_components=16, comm=peer,
                         row_offset=a0, m0_global=X64.shape[0])
    if rank == 0:
        U0, s0, V0 = standard_svd_ref(X64, 16)
        out["standard_fp64"] = {"sigma_rel_err": sigma_rel_err(s.cpu().numpy(), s0),
                                "angle_V_max": float(vector_angles(V.cpu().numpy().T, V0.T).max()),
                                "angle_U_shard0_max": float(vector_angles(U.cpu().numpy(), U0[a0:a1]).max())}
        out["standard_fp64"]["pass"] = bool(out["standard_fp64"]["sigma_rel_err"] < 1e-6 and out["standard_fp64"]["angle_V_max"] < 1e-5)
    # BOP-DMD: trials partitioned over the ranks must reproduce the single-rank ensemble
    from dmd_era5_b200.bopdmd import bopdmd_device

    rng = np.random.RandomState(1)
    om = np.sort(rng.uniform(0.2, 3.0, 4)) + 0.15 * np.arange(4)
    al = np.concatenate([-rng.uniform(0, 0.05, 4) + 1j * om, -rng.uniform(0, 0.05, 4) - 1j * om])
    al[4:] = al[:4].conj()
    Bh = rng.standard_normal((4, 10)) + 1j * rng.standard_normal((4, 10))
    tt = np.linspace(0, 20, 400)
    Hh = (np.exp(np.outer(tt, al)) @ np.concatenate([Bh, Bh.conj()])).real + 0.05 * rng.standard_normal((400, 10))
    sharded = bopdmd_device(ops, Hh, tt, n_trials=41, trial_size=320, r=8, seed=3, comm=comm)
    if rank == 0:
        single = bopdmd_device(ops, Hh, tt, n_trials=41, trial_size=320, r=8, seed=3)
        out["bopdmd_trials_sharded"] = {
            "alphas_max_diff": float((sharded["alphas"] - single["alphas"]).abs().max()),
            "alpha_std_max_diff": float((sharded["alpha_std"] - single["alpha_std"]).abs().max()),
            "mode_mean_max_diff": float((sharded["mode_mean"] - single["mode_mean"]).abs().max())}
        out["bopdmd_trials_sharded"]["pass"] = bool(max(out["bopdmd_trials_sharded"].values()) < 1e-12)
    peer.close()
    if rank == 0:
        out["world"] = world
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
