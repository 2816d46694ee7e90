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


@pytest.mark.timeout(900)
def test_baseline_configs_sharded(tmp_path):
    """BASELINE configs 3, 4, 5 at full size over all GPUs of the box against oracle rows (mgpu_configs_worker.py)."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    worker = Path(__file__).resolve().parent / "mgpu_configs_worker.py"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29537", str(worker), str(tmp_path), "c3,c4,c5", "--p2p"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=850)
    res = (tmp_path / "result.txt").read_text() if (tmp_path / "result.txt").exists() else ""
    assert r.returncode == 0 and res.startswith("OK"), f"{res}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"


def test_device_keyword_selects_a_non_current_gpu(tmp_path):
    """Single process, `device=` pointing at a GPU that is not the current one: same bits as on GPU 0."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import nbody_streams_b200 as nb

    xv, m = nb.make_plummer_sphere(6001, M_total=1e9, a=1.0, seed=4)
    torch.cuda.set_device(0)
    a0, p0 = nb.compute_nbody_forces_potential_gpu(xv[:, :3], m, 0.05)
    a1, p1 = nb.compute_nbody_forces_potential_gpu(xv[:, :3], m, 0.05, device=torch.device("cuda", 1))
    assert np.array_equal(a0, a1) and np.array_equal(p0, p1)
    d1 = nb.compute_nbody_forces_gpu(xv[:, :3], m, 0.05, device=torch.device("cuda", 1), return_cupy=True)
    assert d1.device.index == 1 and np.array_equal(d1.get(), a0)
    kw = dict(precision="float64", kernel="spline", save_snapshots=False, verbose=False)
    r0 = nb.run_nbody_gpu(xv, m, 0.0, 5e-4, 1e-4, 0.05, output_dir=str(tmp_path / "d0"), **kw)
    r1 = nb.run_nbody_gpu(xv, m, 0.0, 5e-4, 1e-4, 0.05, output_dir=str(tmp_path / "d1"), device=torch.device("cuda", 1),
                          **kw)
    assert np.array_equal(r0, r1)
    from nbody_streams_b200.unbind import iterative_unbinding

    (b0,), _, _ = iterative_unbinding(xv[:, :3], 2.0 * xv[:, 3:], m, potential_compute_method="direct_gpu",
                                      verbose=False)
    (b1,), _, _ = iterative_unbinding(xv[:, :3], 2.0 * xv[:, 3:], m, potential_compute_method="direct_gpu",
                                      verbose=False, device="cuda:1")
    assert np.array_equal(b0, b1)
    assert torch.cuda.current_device() == 0
