"""The sharded step on real GPUs (needs >= 2; skipped on a 1-GPU box): the pushed exchange over peer memory, the NCCL
all-gather exchange and the unsharded index give bit-identical fused AND reranked lists over several steps (both buffer
halves, growing sequence numbers, a smaller batch in between).  Runs scripts/exchange_check.py under torchrun."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_pushed_exchange_equals_nccl_equals_unsharded():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("PYTEST_CURRENT_TEST", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", "scripts/exchange_check.py"],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
