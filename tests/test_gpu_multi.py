"""Multi-GPU test: torchrun with one rank per GPU (needs >= 2 GPUs; skipped otherwise)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
def test_sharded_run_matches_single_gpu(tmp_path):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    worker = Path(__file__).resolve().parent / "mgpu_worker.py"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(worker), str(tmp_path), "--p2p"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=550)
    res = (tmp_path / "result.txt").read_text() if (tmp_path / "result.txt").exists() else ""
    assert r.returncode == 0 and res.startswith("OK"), f"{res}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
