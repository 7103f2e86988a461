"""1-GPU result == G-GPU result, bit for bit (needs >= 2 GPUs; skipped otherwise)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_pipeline_equals_single_gpu():
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(ROOT / "tests" / "dist_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dist_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
