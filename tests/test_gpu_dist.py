"""N > 1 parity as a -m gpu test: when the box shows at least two GPUs, the multi-GPU path (samples sharded over ranks,
one NCCL all-to-all-v of bucket regions, partitions sharded over ranks) runs under torchrun at world 2 and at every
visible power of two up to 8, and rank 0 compares every matrix / merge_info / counts file / .pinfo with the CPU oracle
(tests/dist_check.py; hash:bf, kmer:count, k=63 kmer:pa + rescue).  Skipped on a one-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_path_equals_oracle(world):
    if _ngpu() < world:
        pytest.skip(f"{_ngpu()} GPU(s) visible, world {world} needs {world}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" OK") == 3, r.stdout[-2000:]
