"""Multi-rank tests.  The world_size-2 gloo test runs on CPU and covers the exchange protocol of the sharded registration
(global trimmed quantile from all-reduced digit histograms, exact 128-bit partial sums as 32-bit limbs) against the
single-process oracle; the NCCL test needs two GPUs and checks that the sharded CUDA path is bit-identical to the
unsharded one."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _torchrun(nproc, extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "sharded_worker.py")] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def test_sharded_protocol_gloo_world2():
    r = _torchrun(2, ["--backend", "gloo"], 29571)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_SHARDED_OK" in r.stdout


@pytest.mark.gpu
def test_sharded_registration_all_gpus():
    """One registration sharded over every GPU of the box (2, 4 or 8 ranks), both carriers of the exchange, against the
    unsharded registration: bit-identical transform, iteration count, used-point ratio and output cloud."""
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    r = _torchrun(world, ["--backend", "nccl", "--points", "20000"], 29572)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "NCCL_SHARDED_OK" in r.stdout
    assert "peer carrier unavailable" not in r.stdout, r.stdout[-2000:]
