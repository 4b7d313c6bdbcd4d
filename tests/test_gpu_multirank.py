"""GPU, >= 2 devices: the row-sharded paths over NCCL and over the peer-memory communicator (era5svd_comm_*: collectives
against NCCL on random data, the all-reduce fused into the projection's reduction, replicas bit-identical) - randomized
native / tf32x3 / tf32mix, standard FP64, sharded BOP-DMD trials - against the single-process oracles - tests/multigpu_check.py run as a test (multi-rank
NCCL parity was a hand-run script).  Skipped on a one-GPU box; the gloo CPU tests cover the host logic there."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_nccl_parity():
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "multigpu_check.py")], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")][-1])
    assert line["world"] == 2
    for key in ("peer_collectives", "native", "tf32x3", "tf32mix", "peer_native", "peer_tf32mix", "standard_fp64",
                "bopdmd_trials_sharded"):
        assert line[key]["pass"], (key, line[key])


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_comm_falls_back_together_when_one_rank_fails():
    """One rank without peer access (simulated): make_comm must hand NCCL to EVERY rank - no rank may be left inside a
    collective the others never enter - and the peer path must come up again without the failure."""
    port = 29900 + os.getpid() % 90
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "multigpu_fallback_check.py")], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")][-1])
    assert all(line[k] for k in ("fell_back_on_every_rank", "nccl_allreduce_ok", "peer_path_comes_up", "peer_allreduce_ok")), line
