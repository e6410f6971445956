"""GPU, >= 2 devices: the sharded sampler (NCCL all-gather / all-reduce / all-to-all-v) equals the single-GPU run."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_sharded_run_matches_single_gpu():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(ROOT / "tools" / "multi_gpu_check.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
