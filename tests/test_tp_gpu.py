"""Tensor parallelism on real GPUs: tools/tp_parity.py under torchrun with 2, 4 and 8 ranks (each case skips when the box has
fewer devices) -- ids of every variant against the oracle's TP emulation, cached-graph replay across ragged requests, and the
deterministic early exit (all ranks leave the decode loop at the same step)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tensor_parallel_parity(cuda, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "tp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + "\n" + r.stderr)[-6000:]
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, f"tp_parity_{world}.log"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-20000:])
    assert r.returncode == 0 and "TP PARITY PASS" in r.stdout, tail
