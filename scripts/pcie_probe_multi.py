"""Why does the host-to-device leg not scale with the number of GPUs of one node?  

    torchrun --nproc-per-node N scripts/pcie_probe_multi.py

Per rank: pinned-host -> device copy rate of a 1 GiB buffer (a) alone, rank after rank, (b) all ranks at once, each with
its host buffer first-touched (i) wherever the launcher put the process, (ii) after binding the process to the CPUs NVML
names as local to its GPU.  Rank 0 also prints the topology the box reports."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench


def copy_rate(host, dev, reps=4):
    dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return host.numel() * host.element_size() * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "cpus_allowed": len(os.sched_getaffinity(0))}
    if rank == 0:
        for cmd in (["nvidia-smi", "topo", "-m"], ["numactl", "-H"], ["cat", "/sys/fs/cgroup/cpuset.cpus.effective"],
                    ["cat", "/sys/fs/cgroup/cpuset.mems.effective"], ["lscpu"]):
            try:
                txt = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
                print(f"$ {' '.join(cmd)}\n{txt[:3000]}", file=sys.stderr)
            except Exception as e:
                print(f"$ {' '.join(cmd)}: {e!r}", file=sys.stderr)
    dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    for label, bind in (("unbound", False), ("bound to the GPU's CPUs (NVML)", True)):
        info = bench.bind_to_gpu_numa_node(local) if bind else None
        host = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); host.fill_(1)     # first touch here
        alone = torch.zeros(world, device="cuda", dtype=torch.float64)
        for r in range(world):                       # one rank at a time
            dist.barrier()
            if r == rank:
                alone[r] = copy_rate(host, dev)
            dist.barrier()
        dist.all_reduce(alone)
        dist.barrier()
        together = torch.zeros(world, device="cuda", dtype=torch.float64)
        together[rank] = copy_rate(host, dev, reps=8)
        dist.all_reduce(together)
        infos = [None] * world
        dist.all_gather_object(infos, info)
        out[label] = {"alone_GBps": [round(float(x), 1) for x in alone], "together_GBps": [round(float(x), 1) for x in together],
                      "together_sum_GBps": round(float(together.sum()), 1), "binding": infos}
        del host
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
