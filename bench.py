#!/usr/bin/env python
"""Benchmark of the DMD-ERA5 SVD stage: snapshot-matrix GB/s factorised (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c2|c3] [--precision native|tf32x3] [--rows R]

One "step" = one pass of the hot path over one synthetic ERA5-shaped input: matrix build
(stack + time-mean removal + transpose into the space x time snapshot matrix) followed by the
randomized SVD (k = 100, l = 110, n_iter by sklearn's 'auto' rule) through U, s, V on the device.

Workloads (per rank; rows shard over ranks with no data-path collective other than the small
n x l / l x l all-reduces, so N > 1 is WEAK scaling: every rank holds one such shard):
  c2 : 721 x 1440 points x 744 hourly snapshots, float32  -> 1 038 240 x 744  (BASELINE configs[1])
  c3 : 1/8 of 3 vars x 13 levels x 721 x 1440 x 1460 6-hourly, float32 -> 5 061 420 x 1460 per rank
       (8 ranks = BASELINE configs[2] exactly)

value  : whole-job GB/s with the native arrays already resident in HBM (CUDA events, max over ranks)
e2e    : same metric through the host-facing call: pinned host arrays -> H2D -> build -> SVD -> D2H
         of U, s, V inside the timed region
roofline / cpu_baseline : see DESIGN.md "Measurement"
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (points per rank, snapshots, description)
    "c2": (721 * 1440, 744, "0.25deg single level 721x1440 x 744 hourly, f32, randomized k=100 (n_iter=4)"),
    "c3": (3 * 13 * 721 * 1440 // 8, 1460, "1/8 of 0.25deg t/u/v x 13 levels x 1460 6-hourly per rank, f32, randomized k=100 (n_iter=7)"),
}
K_COMPONENTS = 100
METRIC = "snapshot_matrix_GBps_factorised"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clocks / clock-event (throttle) reasons DURING the timed region.  NVML is polled from a thread every
    ~5 ms (nvidia_ml_py: the library `nvidia-smi` itself sits on); `nvidia-smi -lms` is the fallback when NVML cannot
    be loaded — it needs up to a second before its first line, too slow for a 0.3 s timed region on its own."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, uuid: str | None = None):
        self.index = index
        self.uuid = uuid
        self.samples = []                    # (host time, sm MHz, max sm MHz, set of reasons)
        self.proc = None
        self.thread = None
        self.source = None
        self._stop = threading.Event()

    # ---- NVML ----
    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        if self.uuid:
            for u in (self.uuid, "GPU-" + self.uuid):
                try:
                    return pynvml, pynvml.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                except Exception:
                    pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            ent = vis.split(",")[self.index].strip()
            if ent.isdigit():
                idx = int(ent)
            else:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(ent.encode())
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll_nvml(self, pynvml, h):
        bits = [(pynvml.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (pynvml.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        try:
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self._stop.is_set():
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.samples.append((time.perf_counter(), sm, mx, {nm for b, nm in bits if mask & b}))
            except Exception:
                pass
            self._stop.wait(0.005)

    # ---- nvidia-smi fallback ----
    def _poll_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 6:
                continue
            try:
                sm, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            self.samples.append((time.perf_counter(), sm, mx,
                                 {nm for nm, val in zip(self.NAMES, f[2:6]) if val.lower().startswith("active")}))

    def start(self):
        try:
            pynvml, h = self._nvml_handle()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._poll_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self, t0: float | None = None, t1: float | None = None) -> dict:
        """Summary of the samples taken inside [t0, t1] (host clock), i.e. DURING the timed region."""
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        self.thread.join(timeout=2)
        inside = [x for x in self.samples if (t0 is None or x[0] >= t0) and (t1 is None or x[0] <= t1)]
        where = "timed region"
        if not inside:                       # very short region: fall back to the samples nearest to it
            inside, where = self.samples[-3:], "nearest to the timed region"
        sm = [x[1] for x in inside]
        mx = [x[2] for x in inside if x[2] is not None]
        reasons = set().union(*[x[3] for x in inside]) if inside else set()
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source, "window": where}


def bind_to_gpu_numa_node(local_rank: int):
    """Multi-rank runs: keep this rank's threads (and therefore its first-touched pinned host buffers) on the NUMA node
    the GPU hangs off, so that the H2D / D2H copies of the e2e leg do not cross the socket interconnect.  Best effort:
    any failure (no sysfs, restricted cpuset) leaves the affinity untouched."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(local_rank)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if len(allowed) >= 2:
            os.sched_setaffinity(0, allowed)
            return {"node": node, "cpus": len(allowed)}
    except Exception:
        pass
    return None


def cpu_reference_run(rows: int, T: int, k: int, seed: int, threads: int):
    """The reference's own numerics on the host cores for a bounded row sample of the workload:
    NumPy restatement of the build (oracle) + sklearn.utils.extmath.randomized_svd, exactly the
    call of src/dmd_era5/era5_svd/era5_svd.py:258.  Returns (seconds, bytes of the matrix)."""
    from oracle.slice_tools_np import build_matrix_np
    from oracle.svd_ref import randomized_svd_ref

    rng = np.random.RandomState(seed)
    r = 160
    Bt = rng.standard_normal((T, r)).astype(np.float32)
    Bt -= Bt.mean(axis=0)
    Bt = np.linalg.qr(Bt)[0] * (100.0 * 0.93 ** np.arange(r)).astype(np.float32)
    A = (rng.standard_normal((r, rows)) / np.sqrt(rows)).astype(np.float32)
    field = (Bt @ A + 250.0).astype(np.float32)            # (T, rows) native layout
    field = field.reshape(T, 1, 1, rows)
    t0 = time.perf_counter()
    X, _, _ = build_matrix_np([field], True, False, 1)
    U, s, V = randomized_svd_ref(X, k, 1)
    dt = time.perf_counter() - t0
    return dt, X.nbytes, float(s[0])


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, T, desc = WORKLOADS[args.workload]
    threads = len(os.sched_getaffinity(0))
    # bounded sample: ~28 us of host time per row at n = 744 (measured, 16 cores); keep the whole
    # --steps/--warmup run within ~2.5 minutes
    per_step_s = 150.0 / max(1, args.steps + args.warmup)
    rows = args.cpu_rows or int(min(S, 262144, max(32768, per_step_s / 28e-6 * 744 / T)))
    times = []
    for i in range(args.warmup + args.steps):
        dt, nbytes, _ = cpu_reference_run(rows, T, K_COMPONENTS, seed=i, threads=threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = nbytes / 1e9 / (ms / 1e3)
    sample = f"{rows} of {S} rows x {T} snapshots f32 per step (bounded CPU sample of {args.workload})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}", "k": K_COMPONENTS},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("ERA5SVD_PRECISION", "tf32x3"), choices=["native", "tf32x3"],
                    help="tf32x3 (default, headline): tcgen05 3xTF32 passes; native: FP32 FMA passes on the CUDA cores")
    ap.add_argument("--tc-split", default="onchip", choices=["onchip", "hbm"])
    ap.add_argument("--rows", type=int, default=0, help="override points per rank (debug only; invalidates the number)")
    ap.add_argument("--cpu-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from dmd_era5_b200 import _cabi
    from dmd_era5_b200.device_ops import CudaOps, KernelTimer
    from dmd_era5_b200.dist import LocalComm, TorchDistComm
    from dmd_era5_b200.pipeline import build_matrix_device, svd_device
    from dmd_era5_b200.rsvd import n_iter_auto
    from dmd_era5_b200.synthetic import synthetic_field

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
        comm = TorchDistComm()
    else:
        comm = LocalComm()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    S, T, desc = WORKLOADS[args.workload]
    if args.rows:
        S = args.rows
    k = K_COMPONENTS
    ops = CudaOps(device)
    sampler = ClockSampler(local_rank, uuid=str(torch.cuda.get_device_properties(local_rank).uuid))
    if rank == 0:
        sampler.start()
    field = synthetic_field(T, S, device=device, seed=1000 + rank)          # native (T, S) f32, outside the timed region
    m_global = S * world
    row_offset = rank * S
    q = n_iter_auto(m_global, T, k)
    x_bytes = float(S) * T * 4

    def step(src_dev, timer=None):
        ops.timer = timer
        # tf32x3: "onchip" reads the plain matrix and splits it on chip (gemm_tc2.cu); "hbm" streams
        # pre-split hi / lo images written by the build kernel (gemm_tc.cu)
        tc = args.precision == "tf32x3" and args.tc_split == "hbm"
        built = build_matrix_device(ops, [src_dev], mean_center=True, scale=False, split=tc, keep_x=not tc)
        U, s, V = svd_device(ops, built.X, svd_type="randomized", n_components=k, seed=1, precision=args.precision,
                             comm=comm, row_offset=row_offset, m0_global=m_global,
                             split=(built.Xhi, built.Xlo) if tc else None)
        ops.timer = None
        return U, s, V

    def sync_all():
        torch.cuda.synchronize(device)
        comm.barrier()
        torch.cuda.synchronize(device)

    # ---------------- device-resident timing ("value") ----------------
    for _ in range(args.warmup):
        step(field)
    sync_all()
    timer = KernelTimer()
    t_region0 = time.perf_counter()
    launches0 = _cabi.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    # per-launch CUDA events (KernelTimer) ride on every 4th timed step: two event records around each of the ~13
    # timed ops cost ~0.4 ms of a 14 ms step when every step carries them
    sampled = [i for i in range(args.steps) if i % 4 == 0]
    for i in range(args.steps):
        U, s, V = step(field, timer if i % 4 == 0 else None)
    e1.record()
    sync_all()
    launches = _cabi.launch_count() - launches0
    clocks = sampler.stop(t_region0, time.perf_counter()) if rank == 0 else None
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = x_bytes * world / 1e9 / (ms / 1e3)
    ksum = timer.summary()
    s_first = float(s[0].item())

    # ---------------- end-to-end through the host-facing path ----------------
    # Every step's input starts in pinned HOST memory and every step's U, s, V end in pinned host memory.
    #   e2e.value       : SvdStageStream (dmd_era5_b200/stage_stream.py), the public call for a sequence of slices:
    #                     H2D of slice i+1 | build + SVD of slice i | D2H of slice i-1 on three streams
    #   e2e.single_shot : the same copies and compute strictly one after the other (latency of ONE slice)
    e2e = None
    if not args.no_e2e:
        from dmd_era5_b200.stage_stream import SvdStageStream

        host = torch.empty((T, S), dtype=torch.float32, pin_memory=True)
        host.copy_(field)
        stream = SvdStageStream(ops, T, S, n_components=k, svd_type="randomized", precision=args.precision,
                                mean_center=True, scale=False, seed=1, comm=comm, row_offset=row_offset,
                                m0_global=m_global)
        n_e2e = max(3, min(args.steps, 10))
        sig = []
        stream.run([host] * 2)                                           # warm-up (allocations, first-touch)
        sync_all()
        t0 = time.perf_counter()
        stream.run([host] * n_e2e, consume=lambda i, U, s, V: sig.append(float(s[0])))
        sync_all()
        ms_e2e = (time.perf_counter() - t0) * 1e3 / n_e2e
        # one slice, nothing overlapped
        hU, hs, hV = stream.host_out[0]
        dev_in = stream.dev_in[0]

        def single():
            dev_in.copy_(host, non_blocking=True)
            U, s, V = step(dev_in)
            hU.copy_(U, non_blocking=True); hs.copy_(s, non_blocking=True); hV.copy_(V, non_blocking=True)

        single(); sync_all()
        t0 = time.perf_counter()
        for _ in range(3):
            single()
        sync_all()
        ms_single = (time.perf_counter() - t0) * 1e3 / 3
        t = torch.tensor([ms_e2e, ms_single], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e, ms_single = float(t[0].item()), float(t[1].item())
        assert all(abs(x - s_first) <= 1e-6 * s_first for x in sig), "pipelined results differ from the one-shot path"
        e2e = {"value": x_bytes * world / 1e9 / (ms_e2e / 1e3), "unit": "GB/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(stream.h2d_bytes), "d2h_bytes_per_step": int(stream.d2h_bytes),
               "mode": f"SvdStageStream over {n_e2e} host slices: H2D(i+1) | build+SVD(i) | D2H(i-1), 2 device input buffers",
               "single_shot": {"value": x_bytes * world / 1e9 / (ms_single / 1e3), "ms_per_step": ms_single},
               "numa_binding": numa}
        del host, stream

    # ---------------- roofline of the dominant kernel ----------------
    pk = peaks()
    tall = ("sketch", "project", "sketch_tc", "project_tc")
    dom = max((n for n in ksum if n in tall), key=lambda n: ksum[n]["ms"], default=None)
    roofline = None
    if dom:
        d = ksum[dom]
        avg_ms = d["ms"] / d["calls"]
        gbs = d["bytes"] / d["calls"] / (avg_ms / 1e3) / 1e9                 # algorithmic bytes: X read once + tall factor
        t_hbm = d["bytes"] / d["calls"] / (pk["hbm_gbs"] * 1e9)
        if args.precision == "tf32x3":
            # the slower of the two rooflines bounds the pass (BASELINE.json north_star): 3xTF32 issues 3 * 2mnl
            # tensor flops; dense TF32 peak = 1/2 of the measured sustained bf16 peak (kernel timed inside a step)
            peak_tf32 = pk["bf16_tflops"] / 2.0
            # tensor-core products per k-step: 3 (hi*hi + lo*hi + hi*lo); the on-chip-split sketch runs on a
            # tf32-exact small factor (Omega_lo = 0) and issues 2
            nprod = 2.0 if (dom == "sketch_tc" and args.tc_split == "onchip") else 3.0
            tfl = nprod * d["flops"] / d["calls"] / (avg_ms / 1e3) / 1e12
            t_tensor = nprod * d["flops"] / d["calls"] / (peak_tf32 * 1e12)
            if t_tensor >= t_hbm:
                roofline = {"kernel": dom, "bound": "tensor", "achieved": tfl, "peak": peak_tf32, "unit": "TFLOP/s",
                            "frac": tfl / peak_tf32, "traffic": None}
            else:
                roofline = {"kernel": dom, "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": gbs / pk["hbm_gbs"], "traffic": None}
            roofline["note"] = (f"{int(nprod)}xTF32 tcgen05 pass; tensor: {int(nprod)}*2mnl flops vs 1/2 sustained bf16 peak; hbm: algorithmic "
                                f"m*n*4 + m*l*4 bytes (tc_split=hbm: the pre-split hi/lo images double the real X traffic); {pk['source']}")
            roofline["hbm_frac_algorithmic"] = gbs / pk["hbm_gbs"]
            roofline["tensor_frac"] = tfl / peak_tf32
        else:
            roofline = {"kernel": dom, "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / pk["hbm_gbs"], "traffic": None,
                        "note": f"FP32-FMA (CUDA-core) pass, algorithmic bytes m*n*4 + m*l*4 per launch; {pk['source']}"}
        roofline["avg_launch_ms"] = avg_ms
        try:   # DRAM bytes per launch of this kernel from the committed ncu --set full capture (same workload)
            with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
                roofline["traffic"] = json.load(f).get(args.workload, {}).get(dom)
        except Exception:
            pass
    # the other tall pass, for the record: the two-product sketch is HBM bound (algorithmic bytes over the measured copy rate)
    roofline_other = None
    other = {"sketch_tc": "project_tc", "project_tc": "sketch_tc", "sketch": "project", "project": "sketch"}.get(dom)
    if other in ksum and ksum[other]["calls"]:
        o = ksum[other]
        o_ms = o["ms"] / o["calls"]
        o_gbs = o["bytes"] / o["calls"] / (o_ms / 1e3) / 1e9
        roofline_other = {"kernel": other, "bound": "hbm", "achieved": o_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": o_gbs / pk["hbm_gbs"], "avg_launch_ms": o_ms}
        if args.precision == "tf32x3":
            npo = 2.0 if (other == "sketch_tc" and args.tc_split == "onchip") else 3.0
            roofline_other["tensor_frac"] = npo * o["flops"] / o["calls"] / (o_ms / 1e3) / 1e12 / (pk["bf16_tflops"] / 2.0)
            if roofline_other["tensor_frac"] > roofline_other["frac"]:
                roofline_other.update(bound="tensor", achieved=roofline_other["tensor_frac"] * pk["bf16_tflops"] / 2.0,
                                      peak=pk["bf16_tflops"] / 2.0, unit="TFLOP/s", frac=roofline_other["tensor_frac"])
    kernels = {n: {"calls_per_step": v["calls"] / len(sampled), "ms_per_step": v["ms"] / len(sampled)} for n, v in ksum.items()}
    if roofline is not None:
        roofline["timed_launches"] = f"per-launch CUDA events on {len(sampled)} of the {args.steps} timed steps (every 4th)"

    # ---------------- CPU baseline beside it (rank 0, N = 1) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        rows = args.cpu_rows or min(S, 262144)
        dt, nbytes, _ = cpu_reference_run(rows, T, k, seed=0, threads=threads)
        cpu = {"value": nbytes / 1e9 / dt, "unit": "GB/s", "cores": threads, "kind": "reference",
               "sample": f"sklearn randomized_svd + NumPy build on {rows} of {S} rows x {T} f32 ({dt:.2f} s, single run)"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "native" else "tf32x3", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "rows_per_rank": S, "snapshots": T, "k": k, "l": k + 10,
                       "n_iter": q, "mean_center": True, "precision": args.precision, "tc_split": args.tc_split,
                       "l2": "inputs larger than L2 (matrix shard >= 3 GB vs 126 MB)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "roofline_other_pass": roofline_other, "cpu_baseline": cpu,
            "kernels": kernels, "sigma_1": s_first,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
