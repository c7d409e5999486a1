"""-m gpu: the multi-GPU paths on real peers.  When the box has more than one GPU, tests/multi_gpu_check.py is spawned under
torch.distributed.run on min(device_count, 8) of them: head-parallel and sequence-parallel (f16, q8_0; NCCL all-gather, peer-memory
scatter, fused one-kernel step) results are compared with the CPU oracle on every rank.  One GPU: an explicit skip (the same
protocols run at world size 2 over gloo in tests/test_sharding.py, and with emulated ranks in tests/test_decode_gpu.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_paths_match_the_oracle_on_every_rank():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"{n} GPU visible: the multi-rank check needs at least 2")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert f"multi_gpu_check ok on {world} GPUs" in res.stdout
