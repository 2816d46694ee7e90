#!/usr/bin/env python3
"""Per-step time of run_nbody_gpu for BASELINE config 4 (N=262144, float64) and config 3's size (N=1M,
float32_kahan spline) on 1..8 GPUs:  torchrun --nproc-per-node G tools/step_rate_mgpu.py

Per-step time = (t(3S) - t(S)) / 2S (removes set-up, NCCL communicator creation, the initial force evaluation
and the final gather + restart.npz write), wall clock on rank 0 with a barrier on both sides."""
import os
import sys
import tempfile
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nbody_streams_b200 as nb  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def timed(xv, m, steps, h, kw):
    with tempfile.TemporaryDirectory() as td:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nb.run_nbody_gpu(xv, m, 0.0, steps * 1e-4, 1e-4, h, output_dir=td, save_snapshots=False, verbose=False, **kw)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return time.perf_counter() - t0


if rank == 0:
    print(f"| GPUs | N | precision | kernel | S | ms/step | Gint/s aggregate |\n|---:|---:|---|---|---:|---:|---:|")
for N, prec, kernel, h, S in ((262_144, "float64", "plummer", 0.02, 40), (262_144, "float64", "spline", 0.02, 40),
                              (1_000_000, "float32_kahan", "spline", 0.01, 10)):
    xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
    kw = dict(precision=prec, kernel=kernel)
    timed(xv, m, 2, h, kw)                       # warm-up: communicator, module load
    t1 = min(timed(xv, m, S, h, kw) for _ in range(2))
    t3 = min(timed(xv, m, 3 * S, h, kw) for _ in range(2))
    per = (t3 - t1) / (2 * S)
    if rank == 0:
        print(f"| {world} | {N} | {prec} | {kernel} | {S} | {per * 1e3:.3f} | {N * N / per / 1e9:.0f} |", flush=True)
if world > 1:
    dist.destroy_process_group()
