"""Data-parallel parity on real GPUs: torchrun --nproc-per-node 2 tests/dp_check_multigpu.py (NCCL all-gather of the latents,
bucketed gradient all-reduce, code-book EMA all-reduce) against the single-process step at the global batch: loss, RAW gradients
(before Adam), updated parameters, code-book. Skipped on a single-GPU box."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run with gpurun --gpus 2)")
def test_data_parallel_step_matches_single_process_nccl():
    root = Path(__file__).resolve().parents[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(root / "tests" / "dp_check_multigpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(root))
    out = r.stdout + r.stderr
    print(out[-3000:])
    assert r.returncode == 0 and "DP_CHECK PASS" in out, out[-3000:]
