"""The NCCL transport under test (one process per GPU, torchrun): the union of the strips equals
the single engine bit for bit on three scenes (scripts/dist_check.py).  Needs at least two visible
GPUs; the single-GPU CI box runs the same logic through the in-process group (test_strips_gpu.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="the NCCL transport needs two GPUs")
def test_nccl_strips_equal_single_engine():
    n = min(_gpus(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "scripts", "dist_check.py"), "40"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count(" OK") == 3
