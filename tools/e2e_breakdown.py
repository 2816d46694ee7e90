#!/usr/bin/env python3
"""Where does the time of one multi-rank fields-API call go?  torchrun --nproc-per-node G tools/e2e_breakdown.py"""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200 import fields  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 1_000_000
xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
pos = np.ascontiguousarray(xv[:, :3].astype(np.float32))
g = dist.group.WORLD


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(4):
    dist.barrier()
    t0 = sync()
    ops = fields._prepare(pos, float(m[0]), 0.01, "float32_kahan", "plummer", False, None, g, None)
    pos_t, mass_t, mass_s, h_t, h_s, N_, prec, kid, dtn, loc = ops
    t1 = sync()
    acc, phi = fields._run_direct(pos_t, mass_t, mass_s, h_t, h_s, N_, prec, kid, 3, nb.G_DEFAULT, 1e-15, dtn, g, None, None,
                                  "all", loc)
    t2 = sync()
    a, p = fields._to_host(acc), fields._to_host(phi)
    t3 = sync()
    if rank == 0 and it:
        print(f"prepare (H2D) {1e3 * (t1 - t0):.2f} ms | pack+exchange+sweep+reduce-scatter+finish+gather {1e3 * (t2 - t1):.2f} ms | "
              f"D2H {1e3 * (t3 - t2):.2f} ms | total {1e3 * (t3 - t0):.2f} ms", flush=True)
# finer: time the pieces of _run_direct with events by calling the public function under the profiler
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    nb.compute_nbody_forces_potential_gpu(pos, float(m[0]), 0.01, precision="float32_kahan", kernel="plummer", group=g)
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
dist.destroy_process_group()
