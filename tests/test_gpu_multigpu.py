"""``contract_distributed`` (cotengra/core.py:4032-4090 ``contract_mpi``) on real GPUs over NCCL:
N ranks launched with torchrun run ``scripts/gpu_dist_check.py`` -- all-reduce and reduce-to-root,
stripped exponents, sliced-output sharding -- against the golden values of the unmodified
reference.  Needs >= 2 visible GPUs (``gpurun --gpus 2``); skipped on a 1-GPU box."""

import os
import socket
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_contract_distributed_nccl(world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "scripts", "gpu_dist_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stderr[-4000:]
    assert "DIST_CHECK PASS" in res.stdout
