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
    PERIOD = 0.01                            # seconds between NVML polls (every 5 ms an occasional poll delayed a step by ~2 ms, profiles/r02_step_jitter.txt)

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
            self._stop.wait(self.PERIOD)

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
            return {"node": node, "cpus": len(allowed), "source": "sysfs"}
    except Exception:
        pass
    try:                                  # containers often hide the PCI sysfs tree: ask NVML for the GPU's ideal CPUs
        import pynvml
        import torch

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = None
        for u in (uuid, "GPU-" + uuid):
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(u.encode())
                break
            except Exception:
                pass
        if h is None:
            return None
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0) & cpus
        if len(allowed) >= 2:
            os.sched_setaffinity(0, allowed)
            return {"cpus": len(allowed), "of_ideal": len(cpus), "source": "nvml"}
        return {"cpus": 0, "of_ideal": len(cpus), "source": "nvml", "note": "none of the GPU's ideal CPUs is in this process's cpuset"}
    except Exception:
        pass
    return None


def set_blas_threads(n: int):
    """torchrun exports OMP_NUM_THREADS=1 to every rank, which silently made the N > 1 reference arm single threaded
    (VERDICT r01 weak #14).  The CPU legs set the BLAS pool explicitly and report what the pool says."""
    info = []
    try:
        import threadpoolctl

        threadpoolctl.threadpool_limits(limits=n)
        info = [{k: d.get(k) for k in ("user_api", "internal_api", "num_threads", "version")}
                for d in threadpoolctl.threadpool_info()]
    except Exception as e:      # pragma: no cover
        info = [{"error": repr(e)}]
    return info


def sample_field_np(rows: int, T: int, seed: int) -> np.ndarray:
    """Host generator of a (T, rows) float32 native-layout field with the bench spectrum (sigma_i = 100 * 0.93**i,
    160 components, mean 250) - the reference arm's input when no GPU-generated sample is handed over."""
    rng = np.random.RandomState(seed)
    r = 160
    Bt = rng.standard_normal((T, r)).astype(np.float32)
    Bt -= Bt.mean(axis=0)
    Bt = np.linalg.qr(Bt)[0] * (100.0 * 0.93 ** np.arange(r)).astype(np.float32)
    A = (rng.standard_normal((r, rows)) / np.sqrt(rows)).astype(np.float32)
    return (Bt @ A + 250.0).astype(np.float32)


def cpu_reference_run(field: np.ndarray, k: int, seed: int):
    """The reference's own numerics on the host cores for a bounded row sample of the workload:
    NumPy restatement of the build (oracle) + sklearn.utils.extmath.randomized_svd, exactly the
    call of src/dmd_era5/era5_svd/era5_svd.py:258.  field: (T, rows) native layout.
    Returns (seconds, bytes of the matrix, X, (U, s, V))."""
    from oracle.slice_tools_np import build_matrix_np
    from oracle.svd_ref import randomized_svd_ref

    T, rows = field.shape
    t0 = time.perf_counter()
    X, _, _ = build_matrix_np([field.reshape(T, 1, 1, rows)], True, False, 1)
    U, s, V = randomized_svd_ref(X, k, seed)
    dt = time.perf_counter() - t0
    return dt, X.nbytes, X, (U, s, V)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, T, desc = WORKLOADS[args.workload]
    threads = len(os.sched_getaffinity(0))
    pool = set_blas_threads(threads)
    # bounded sample: ~28 us of host time per row at n = 744 (measured, 16 cores); keep the whole
    # --steps/--warmup run within ~2.5 minutes
    per_step_s = 150.0 / max(1, args.steps + args.warmup)
    rows = args.cpu_rows or int(min(S, 262144, max(32768, per_step_s / 28e-6 * 744 / T)))
    times = []
    for i in range(args.warmup + args.steps):
        dt, nbytes, _, _ = cpu_reference_run(sample_field_np(rows, T, seed=i), K_COMPONENTS, seed=1)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = nbytes / 1e9 / (ms / 1e3)
    sample = (f"{rows} of {S} rows x {T} snapshots f32 per step (bounded CPU sample of {args.workload}; GB/s is per byte "
              f"of the sampled matrix, every phase of the reference is O(rows))")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "rows_per_rank": S, "snapshots": T, "k": K_COMPONENTS,
                   "l": K_COMPONENTS + 10, "n_iter": 7 if K_COMPONENTS < 0.1 * min(S, T) else 4, "mean_center": True,
                   "sampled_rows": rows, "same_config_as_gpu_arm": False,
                   "note": "the reference cannot hold / finish the full matrix in bench time: a row sample, GB/s-normalised"},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "reference", "sample": sample,
                         "threadpool_info": pool, "OMP_NUM_THREADS_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# tensor-core products per k-step of each timed tall kernel (KernelTimer names, device_ops.py)
NPROD = {"sketch_x1": 1.0, "project_x1": 1.0, "sketch_tc": 2.0, "project_x2": 2.0, "project_tc": 3.0, "gram_tc": 3.0,
         "apply_basis_tc": 3.0}
TALL = ("sketch", "project", "sketch_tc", "project_tc", "sketch_x1", "project_x1", "project_x2")


def pass_rooflines(ksum: dict, pk: dict, precision: str, tc_split: str) -> list[dict]:
    """One entry per timed tall kernel: achieved GB/s over the ALGORITHMIC bytes (X once + the tall factor), tensor
    TFLOP/s over the products actually issued, and frac = max(t_hbm, t_tensor) / t_kernel against the MEASURED peaks
    (HBM copy rate from MEASURED_PEAKS.json, kind::tf32 rate from this run's own probe)."""
    out = []
    for name, d in ksum.items():
        if not d["calls"] or (name not in TALL and name != "build_rows"):
            continue
        ms = d["ms"] / d["calls"]
        nbytes, flops = d["bytes"] / d["calls"], d["flops"] / d["calls"]
        t_hbm = nbytes / (pk["hbm_gbs"] * 1e9) * 1e3
        e = {"kernel": name, "calls": d["calls"], "avg_launch_ms": ms, "algorithmic_GB": nbytes / 1e9,
             "achieved_GBps": nbytes / 1e9 / (ms / 1e3), "t_hbm_ms": t_hbm, "hbm_frac": t_hbm / ms}
        t_tensor = 0.0
        if precision != "native" and name in NPROD:
            nprod = NPROD[name]
            if name == "sketch_tc" and tc_split == "hbm":
                nprod = 3.0
            t_tensor = nprod * flops / (pk["tf32_tflops"] * 1e12) * 1e3
            e.update(tensor_products=nprod, achieved_TFLOPs=nprod * flops / 1e12 / (ms / 1e3), t_tensor_ms=t_tensor,
                     tensor_frac=t_tensor / ms)
        e["bound"] = "tensor" if t_tensor > t_hbm else "hbm"
        e["frac"] = max(t_hbm, t_tensor) / ms
        out.append(e)
    return sorted(out, key=lambda e: -e["avg_launch_ms"] * e["calls"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("ERA5SVD_PRECISION", "tf32mix"),
                    choices=["native", "tf32x3", "tf32mix"],
                    help="tf32mix (default, headline): tcgen05 passes, single-product TF32 for the early power iterations, "
                         "3xTF32 for the last iteration + final passes; tf32x3: 3xTF32 everywhere; native: FP32 FMA")
    ap.add_argument("--tc-split", default="onchip", choices=["onchip", "hbm"])
    ap.add_argument("--rows", type=int, default=0, help="override points per rank (debug only; invalidates the number)")
    ap.add_argument("--cpu-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-north-star", action="store_true", help="skip the c3 (north-star shape) leg")
    ap.add_argument("--no-collectives", action="store_true",
                    help="diagnosis only (the results are WRONG and the line says so): every rank factorises its shard "
                         "alone, which isolates the rank skew from the cost of the collectives")
    ap.add_argument("--north-star-steps", type=int, default=5)
    ap.add_argument("--balance", action="store_true",
                    help="north-star leg: second run with row shards proportional to each rank's measured speed (measured at 8 "
                         "GPUs: speeds within +-3.5 %%, 119.5 against 117.8 ms with equal shards - not faster, hence opt-in)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from dmd_era5_b200 import _cabi
    from dmd_era5_b200.device_ops import CudaOps, KernelTimer
    from dmd_era5_b200.dist import LocalComm, PeerComm, make_comm, shard_rows, shard_rows_weighted
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
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    k = K_COMPONENTS
    ops = CudaOps(device)
    # collectives: our kernels over peer-mapped memory (csrc/comm.cu; the all-reduce of Z fused into the projection's
    # reduction) when the GPUs of the node can map each other, else torch.distributed / NCCL; ERA5SVD_COMM=nccl forces NCCL
    comm = make_comm(ops) if world > 1 else LocalComm()
    comm_kind = "none (1 rank)" if world == 1 else ("peer memory (era5svd_comm_*, fused into reduce_partials)"
                                                     if isinstance(comm, PeerComm) else "nccl (torch.distributed)")
    if args.no_collectives and world > 1:
        class _NoCollectives(LocalComm):        # barrier only: each rank's SVD is that of its own shard (diagnosis)
            barrier = staticmethod(comm.barrier)
        comm = _NoCollectives()
        comm_kind = "NONE - DIAGNOSIS RUN, results are per-shard SVDs, not the row-sharded SVD (invalid as a benchmark)"

    def sync_all():
        torch.cuda.synchronize(device)
        comm.barrier()
        torch.cuda.synchronize(device)

    # ---------------- measured peaks (this run, this GPU): roofline denominators ----------------
    import ctypes as C

    pk = peaks()
    probe = {}
    try:
        v = C.c_double(0.0)
        for key, form, N, secs in (("tf32_tflops_burst", 1, 256, 0.02), ("tf32_tflops_sustained", 1, 256, 0.4),
                                   ("tf32_tflops_n112_ts", 1, 112, 0.05), ("tf32_tflops_n112_ss", 0, 112, 0.05)):
            _cabi.check(ops.lib.era5svd_probe_tf32_tflops(form, N, secs, C.byref(v)), "era5svd_probe_tf32_tflops")
            probe[key] = v.value
        _cabi.check(ops.lib.era5svd_probe_dmma_tflops(0.2, C.byref(v)), "era5svd_probe_dmma_tflops")
        probe["fp64_dmma_tflops"] = v.value
    except Exception as e:      # a failed probe must not void the bench: fall back and say so
        probe["error"] = repr(e)
    # a tall pass is timed INSIDE a long step under the power cap -> the sustained figure is the denominator
    pk["tf32_tflops"] = probe.get("tf32_tflops_sustained") or pk["bf16_tflops"] / 2.0
    pk["tf32_source"] = ("measured in this run: era5svd_probe_tf32_tflops (every SM issuing tcgen05.mma kind::tf32 M=128 N=256 K=8 "
                         "from TMEM for 0.4 s)" if "tf32_tflops_sustained" in probe else "fallback: 1/2 of the sustained bf16 peak")
    sync_all()

    def time_workload(field, m_global, row_offset, steps, warmup, precision, sample_clocks):
        """warmup + steps full steps (build + randomized SVD) on the resident native array ``field``."""
        S_loc, T_loc = field.shape[1], field.shape[0]

        def step(src_dev, timer=None):
            ops.timer = timer
            # tf32x3 / tf32mix: "onchip" reads the plain matrix and splits it on chip (gemm_tc2.cu); "hbm" streams
            # pre-split hi / lo images written by the build kernel (gemm_tc.cu)
            tc = precision == "tf32x3" and args.tc_split == "hbm"
            built = build_matrix_device(ops, [src_dev], mean_center=True, scale=False, split=tc, keep_x=not tc)
            U, s, V = svd_device(ops, built.X, svd_type="randomized", n_components=k, seed=1, precision=precision,
                                 comm=comm, row_offset=row_offset, m0_global=m_global,
                                 split=(built.Xhi, built.Xlo) if tc else None)
            ops.timer = None
            return U, s, V

        sampler = None
        if sample_clocks and rank == 0:
            sampler = ClockSampler(local_rank, uuid=str(torch.cuda.get_device_properties(local_rank).uuid))
            sampler.start()
        for _ in range(warmup):
            U, s, V = step(field)                 # results held like in the timed loop: the allocator's steady state
        sync_all()
        timer = KernelTimer()
        t_region0 = time.perf_counter()
        launches0 = _cabi.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        # per-launch CUDA events (KernelTimer) ride on every 4th timed step: two event records around each of the ~13
        # timed ops cost ~0.4 ms of a 14 ms step when every step carries them
        sampled = [i for i in range(steps) if i % 4 == 0]
        step_ev = [] if os.environ.get("ERA5SVD_BENCH_STEP_EVENTS") else None   # diagnosis: one event per step
        host_t, seg0 = [], None
        if step_ev is not None:
            import gc
            seg0 = (torch.cuda.memory_stats(device).get("num_device_alloc", 0), gc.get_count(), [g["collections"] for g in gc.get_stats()])
        for i in range(steps):
            h0 = time.perf_counter()
            U, s, V = step(field, timer if i % 4 == 0 else None)
            if step_ev is not None:
                step_ev.append(torch.cuda.Event(enable_timing=True)); step_ev[-1].record()
                host_t.append(round((time.perf_counter() - h0) * 1e3, 2))
        e1.record()
        sync_all()
        if step_ev:
            per = [round(a.elapsed_time(b), 2) for a, b in zip([e0] + step_ev[:-1], step_ev)]
            print(f"[bench] rank {rank} per-step device ms: {per}", file=sys.stderr)
            print(f"[bench] rank {rank} per-step host enqueue ms: {host_t}; cudaMalloc calls in the region: "
                  f"{torch.cuda.memory_stats(device).get('num_device_alloc', 0) - seg0[0]}; gc collections before / after: "
                  f"{seg0[2]} / {[g['collections'] for g in gc.get_stats()]}", file=sys.stderr)
        launches = _cabi.launch_count() - launches0
        clocks = sampler.stop(t_region0, time.perf_counter()) if sampler is not None else None
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ksum = timer.summary()
        for v_ in ksum.values():                   # per sampled step -> totals stay, add the sample count
            v_["sampled_steps"] = len(sampled)
        return {"ms": float(t.item()), "ksum": ksum, "launches": int(launches), "clocks": clocks,
                "sigma_1": float(s[0].item()), "step": step, "sampled": len(sampled)}

    # ---------------- device-resident timing ("value"): BASELINE configs[1] (c2) per rank, weak scaling ----------
    S, T, desc = WORKLOADS[args.workload]
    if args.rows:
        S = args.rows
    # native (T, S) f32, outside the timed region; the ranks' shards form one field (shared temporal patterns)
    field = synthetic_field(T, S, device=device, seed=1000 + rank, time_seed=1000 if world > 1 else None,
                            total_points=S * world)
    m_global = S * world
    row_offset = rank * S
    q = n_iter_auto(m_global, T, k)
    x_bytes = float(S) * T * 4
    res = time_workload(field, m_global, row_offset, args.steps, args.warmup, args.precision, True)
    ms, ksum, launches, clocks, s_first, step = (res[x] for x in ("ms", "ksum", "launches", "clocks", "sigma_1", "step"))
    value = x_bytes * world / 1e9 / (ms / 1e3)

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
               "mode": f"SvdStageStream over {n_e2e} host slices: H2D(i+1) | build+SVD(i) | D2H(i-1), 2 device input buffers "
                       "(pipelined throughput; single_shot is the latency of one slice with nothing overlapped)",
               "single_shot": {"value": x_bytes * world / 1e9 / (ms_single / 1e3), "ms_per_step": ms_single},
               "pcie_floor_ms": stream.h2d_bytes / 55.6e9 * 1e3,
               "numa_binding": numa}
        del host, stream

    # ---------------- roofline of every tall kernel + the dominant one ----------------
    passes = pass_rooflines(ksum, pk, args.precision, args.tc_split)
    tall_passes = [e for e in passes if e["kernel"] in TALL]
    roofline = None
    if tall_passes:
        dom = tall_passes[0]
        unit_t = dom["bound"] == "tensor"
        roofline = {"kernel": dom["kernel"], "bound": dom["bound"],
                    "achieved": dom["achieved_TFLOPs"] if unit_t else dom["achieved_GBps"],
                    "peak": pk["tf32_tflops"] if unit_t else pk["hbm_gbs"], "unit": "TFLOP/s" if unit_t else "GB/s",
                    "frac": dom["frac"], "traffic": None, "avg_launch_ms": dom["avg_launch_ms"],
                    "t_hbm_ms": dom["t_hbm_ms"], "t_tensor_ms": dom.get("t_tensor_ms"),
                    "hbm_peak_GBps": pk["hbm_gbs"], "tf32_peak_TFLOPs": pk["tf32_tflops"],
                    "note": ("frac = max(t_hbm, t_tensor) / t_kernel: the slower of the two MEASURED rooflines bounds the pass "
                             f"(BASELINE.json north_star); hbm: {pk['source']}; tf32: {pk['tf32_source']}; algorithmic bytes = "
                             "m*n*4 + m*l*4 (+ n*l*8), tensor flops = products per k-step x 2mnl"),
                    "timed_launches": f"per-launch CUDA events on {res['sampled']} of the {args.steps} timed steps (every 4th)"}
        try:   # DRAM bytes per launch of this kernel from a COMMITTED ncu --set full capture (same workload): a pointer,
               # not a measurement of this run
            for fn in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
                pth = os.path.join(ROOT, "profiles", fn)
                if os.path.exists(pth):
                    with open(pth) as f:
                        tr = json.load(f)
                    val = tr.get(args.workload, {}).get(dom["kernel"])
                    if val is not None:
                        roofline["traffic"] = val
                        roofline["traffic_source"] = f"committed ncu capture ({tr.get('_source', fn)}), not measured in this run"
                        break
        except Exception:
            pass
    kernels = {n: {"calls_per_step": v["calls"] / res["sampled"], "ms_per_step": v["ms"] / res["sampled"]}
               for n, v in ksum.items()}

    del field
    torch.cuda.empty_cache()
    ops.reserve_small_pool()

    # ---------------- north-star leg: BASELINE configs[2] (c3: 40 491 360 x 1460, k = 100, n_iter = 7) ----------------
    # N >= 4: the WHOLE matrix, row-sharded (strong) over the N ranks - the north-star number.  N < 4: it does not fit
    # (236.5 GB of X + the native arrays), so one rank's 1/8 share is timed instead (what each rank of the 8-GPU run
    # executes, minus the all-reduces) and labelled as such.
    north = None
    if not args.no_north_star and args.workload == "c2" and not args.rows:
        M3, T3 = 3 * 13 * 721 * 1440, 1460
        if world >= 4:
            r0, r1 = shard_rows(M3, world, rank)
            label = f"c3 WHOLE: {M3} x {T3} f32 (236.5 GB) row-sharded over {world} GPUs (strong), randomized k=100, n_iter=7"
            mg, ro = M3, r0
        else:
            r0, r1 = 0, M3 // 8
            label = (f"c3 SHARD ONLY: 1/8 of the rows ({r1} x {T3} f32, 29.6 GB) on each of {world} GPU(s) - the whole matrix "
                     "needs >= 4 GPUs; this is one rank's share of the 8-GPU run, not the north-star number")
            mg, ro = r1 * world, rank * r1
        # ONE field over all ranks: shared temporal patterns, per-rank spatial patterns (synthetic.py)
        f3 = synthetic_field(T3, r1 - r0, device=device, seed=2000 + rank, time_seed=77, total_points=mg)
        r3 = time_workload(f3, mg, ro, args.north_star_steps, 3, args.precision, True)
        tot_bytes = float(M3 if world >= 4 else (r1 - r0) * world) * T3 * 4
        balance = None
        if world >= 4:
            # Strong scaling of ONE matrix: every collective waits for the slowest rank.  The line reports the speed each rank
            # showed (device time of its own tall kernels per row, per-launch events, the waits excluded); with --balance a
            # second run uses row shards proportional to it and both runs are reported.
            try:
                busy = sum(v["ms"] for nm, v in r3["ksum"].items() if nm in TALL or nm == "build_rows")
                mine = torch.tensor([float(r1 - r0), float(busy)], device=device, dtype=torch.float64)
                every = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(every, mine)
                speed = [float(e[0] / e[1]) for e in every]
                mean_speed = sum(speed) / world
                weights = [min(1.15, max(0.85, sp / mean_speed)) for sp in speed]
                equal = {"ms_per_step": r3["ms"], "GBps": tot_bytes / 1e9 / (r3["ms"] / 1e3), "rows_this_rank": r1 - r0}
                if args.balance and max(speed) / min(speed) > 1.02:
                    del f3
                    torch.cuda.empty_cache()
                    ops.reserve_small_pool()
                    r0, r1 = shard_rows_weighted(M3, weights, rank)
                    ro = r0
                    f3 = synthetic_field(T3, r1 - r0, device=device, seed=2000 + rank, time_seed=77, total_points=mg)
                    r3b = time_workload(f3, mg, ro, args.north_star_steps, 3, args.precision, True)
                    balance = {"sharding": "rows proportional to each rank's measured speed (tall-kernel device time per row in "
                                           "the equal-shard run of this same process)",
                               "relative_speed_per_rank": [round(sp / mean_speed, 4) for sp in speed],
                               "equal_shards": equal, "balanced_ms_per_step": r3b["ms"]}
                    if r3b["ms"] < r3["ms"]:
                        r3 = r3b
                        label += "; row shards proportional to the measured per-rank speed"
                    else:
                        balance["note"] = "the balanced run was not faster: the equal-shard numbers are reported"
                else:
                    balance = {"sharding": "equal", "relative_speed_per_rank": [round(sp / mean_speed, 4) for sp in speed],
                               "note": "speed = rows per ms of each rank's own tall kernels (per-launch events, waits for other "
                                       "ranks excluded); --balance re-runs with shards proportional to it"}
            except Exception as e:                      # noqa: BLE001 - the equal-shard result stands
                balance = {"sharding": "equal", "error": repr(e)[:200]}
        p3 = pass_rooflines(r3["ksum"], pk, args.precision, args.tc_split)
        north = {"workload": label, "whole_matrix": world >= 4, "ms_per_step": r3["ms"],
                 "GBps": tot_bytes / 1e9 / (r3["ms"] / 1e3), "steps": args.north_star_steps, "warmup": 3,
                 "rows_this_rank": r1 - r0, "n_iter": n_iter_auto(mg, T3, k), "sigma_1": r3["sigma_1"], "clocks": r3["clocks"],
                 "passes": [{x: e[x] for x in ("kernel", "calls", "avg_launch_ms", "achieved_GBps", "hbm_frac", "bound", "frac")
                             if x in e} | ({"tensor_frac": e["tensor_frac"], "achieved_TFLOPs": e["achieved_TFLOPs"]}
                                           if "tensor_frac" in e else {}) for e in p3],
                 "load_balance": balance,
                 "min_tall_pass_frac": min((e["frac"] for e in p3 if e["kernel"] in TALL), default=None),
                 "target": ">= 0.60 of the slower of the HBM and tensor rooflines per pass (BASELINE.json north_star)"}
        del f3, r3
        torch.cuda.empty_cache()
        ops.reserve_small_pool()

    # ---------------- CPU baseline + parity IN THE SAME RUN (rank 0, N = 1) ----------------
    # The same <= 262 144-row sample of the workload goes through the reference's own calls on the host cores
    # (cpu_baseline, timed) and through the CUDA path (parity: BASELINE.md section 5).
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.compare import recon_rel_err, sigma_rel_err, vector_angles

        threads = len(os.sched_getaffinity(0))
        pool = set_blas_threads(threads)
        rows = args.cpu_rows or min(S, 262144)
        fs = synthetic_field(T, rows, device=device, seed=77)
        fs_h = fs.cpu().numpy()
        dt, nbytes, Xs, (U0, s0, V0) = cpu_reference_run(fs_h, k, seed=1)
        cpu = {"value": nbytes / 1e9 / dt, "unit": "GB/s", "cores": threads, "kind": "reference",
               "sample": f"sklearn randomized_svd + NumPy build on {rows} of {S} rows x {T} f32 ({dt:.2f} s, single cold run; "
                         "the --impl reference arm reports a warm mean over its steps)",
               "threadpool_info": pool}
        # SVD parity on IDENTICAL input: the matrix the reference path built on the host goes through the CUDA SVD.
        # (The two builds differ in the last float32 digits of the time mean - NumPy accumulates a float32 mean in
        # float32, pairwise; the kernel accumulates in float64 - which on ~250 K data is a per-row offset of up to
        # ~3e-4, i.e. a rank-one artefact of the REFERENCE's matrix; reported separately as build_max_abs_diff.)
        from dmd_era5_b200.era5_svd import host_to_device_matrix

        built = build_matrix_device(ops, [fs], mean_center=True, scale=False)
        build_diff = float(np.max(np.abs(built.X.cpu().numpy() - Xs)))
        mean_diff = float(np.max(np.abs(built.mean.cpu().numpy().astype(np.float64) - fs_h.astype(np.float64).mean(axis=0))))
        Ug, sg, Vg = svd_device(ops, host_to_device_matrix(ops, Xs), svd_type="randomized", n_components=k, seed=1,
                                precision=args.precision)
        Ug, sg, Vg = Ug.cpu().numpy(), sg.cpu().numpy(), Vg.cpu().numpy()
        ang = vector_angles(Ug, U0)
        ref_rec = recon_rel_err(Xs, U0, s0, V0)
        parity = {"sample_rows": rows, "against": "the reference's own float32 call on the same matrix (sklearn randomized_svd, "
                                                 "np.random.seed(1)), era5_svd.py:258",
                  "sigma_rel_err": sigma_rel_err(sg, s0), "max_angle_rad_first50": float(ang[:50].max()),
                  "max_angle_rad": float(ang.max()), "recon_ratio": recon_rel_err(Xs, Ug, sg, Vg) / ref_rec,
                  "build_max_abs_diff": build_diff, "device_mean_vs_float64_mean_max_abs": mean_diff,
                  "build_note": "device build vs NumPy float32 build of the same field; the difference is the float32 "
                                "summation error of NumPy's mean (the device mean is compared with the float64 mean)",
                  "tolerance": "sigma 1e-4 (FP32-split mode), recon within 1 %"}
        del fs, built

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"native": "f32", "tf32x3": "tf32x3", "tf32mix": "tf32x3+tf32"}[args.precision], "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "rows_per_rank": S, "snapshots": T, "k": k, "l": k + 10,
                       "n_iter": q, "mean_center": True, "precision": args.precision, "tc_split": args.tc_split,
                       "precision_note": ("tf32mix: 3xTF32 (fp32-level) products for the last power iteration and the final range / "
                                          "projection passes, single-product TF32 for the earlier power iterations"
                                          if args.precision == "tf32mix" else None),
                       "l2": "inputs larger than L2 (matrix shard >= 3 GB vs 126 MB)", "collectives": comm_kind},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "roofline_passes": passes, "peaks_measured": {"hbm_gbs": pk["hbm_gbs"], "hbm_source": pk["source"], **probe},
            "north_star": north, "cpu_baseline": cpu, "parity": parity,
            "kernels": kernels, "sigma_1": s_first,
        }
        print(json.dumps(out))
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
