"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): the library's one-kernel gradient exchange over peer
memory (include/gscomm_b200.h, P2P and NVLS modes) against NCCL's all_reduce — within float rounding, bit-identical on
every rank, repeated calls — and the keyframe-sharded map step against single-rank accumulation.
The world_size-2 host logic is covered on CPU by tests/test_mapstep_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_exchange_and_sharded_step(built_lib):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multirank_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIRANK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
