"""-m gpu, needs >= 2 GPUs (skipped otherwise): tensor-parallel engine (NCCL all-reduce / all-gather, sharded weights,
shared-memory step plans) vs the CPU oracle, launched one process per GPU with torchrun."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("tp,cfg", [(2, "tiny")])
def test_tensor_parallel_matches_oracle(tp, cfg):
    if torch.cuda.device_count() < tp:
        pytest.skip(f"needs {tp} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={tp}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(ROOT / "scripts" / "tp_check.py"), cfg]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert "TP_CHECK_PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
