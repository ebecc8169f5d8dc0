"""Multi-rank tests: world_size-2/4 gloo on CPU (halo plans + exchange ordering) and, on the GPU
box, NCCL runs whose bricks must equal the single-GPU run bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mp", "halo_worker.py")


def _run(world, mode, bc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, mode, bc]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MP_RESULT bad=0" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("world,bc,port", [(2, "periodic", 29611), (2, "outflow", 29612), (4, "periodic", 29613)])
def test_gloo_halo_exchange_matches_single_process(world, bc, port):
    _run(world, "cpu", bc, port)


@pytest.mark.gpu
@pytest.mark.parametrize("bc,port", [("periodic", 29621), ("outflow", 29622)])
def test_nccl_bricks_match_single_gpu(bc, port):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(2 if n < 4 else 4, "gpu", bc, port)
