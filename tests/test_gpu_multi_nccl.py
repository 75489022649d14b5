"""The multi-GPU paths over real NCCL, inside pytest: needs >= 2 GPUs, skipped on a single-GPU box.

Both checks launch one process per GPU with torchrun on 127.0.0.1 and compare against a single-GPU run of the same
system: the Morton-slice driver (bh_mg_*, C++/NCCL behind the C ABI) must be bit-identical, the locally-essential-tree
exchange must agree to float noise (its summation order differs: own tree + ghost tree).  The emulated-rank versions
of both (one GPU, no NCCL) are tests/test_gpu_parity.py::test_two_morton_slices... and tests/test_gpu_let.py.
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script, world, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", script)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stdout + r.stderr


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_morton_slices_over_nccl_equal_one_gpu():
    world = min(_gpus(), 8)
    world = 1 << (world.bit_length() - 1)   # 2, 4 or 8
    rc, out = _torchrun("check_sliced.py", world)
    assert rc == 0 and f"SLICED_CHECK PASS world {world}" in out, out[-3000:]


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_locally_essential_tree_over_nccl_equals_one_gpu():
    rc, out = _torchrun("check_let.py", 2)
    assert rc == 0 and "LET_CHECK PASS world 2" in out, out[-3000:]
