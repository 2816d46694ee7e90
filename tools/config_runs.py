#!/usr/bin/env python3
"""BASELINE configurations 4 and 5 as actual runs (1 GPU, or N GPUs under torchrun):

  C4  Plummer N=262144, softening 0.02, float64, KDK dt=1e-4, energy / momentum every 50 steps
  C5  dark 1.6M + star 400k + BH (N=2,000,001), float32, spline (run_simulation), a few steps

    python tools/config_runs.py [--c4-steps 1000] [--c5-steps 5]
    torchrun --nproc-per-node 8 tools/config_runs.py ...
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = [f"# BASELINE configs 4 and 5 as runs on {world} GPU(s) (round 2: pair-symmetric sweep)\n"]

steps = arg("--c4-steps", 1000)
if steps > 0:
    N = 262_144
    xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
    for kernel in ("plummer",):
        diags = []
        torch.cuda.synchronize(); t0 = time.perf_counter()
        nb.run_nbody_gpu(xv, m, 0.0, steps * 1e-4, 1e-4, 0.02, precision="float64", kernel=kernel,
                         save_snapshots=False, verbose=False, output_dir=f"/tmp/c4_{rank}", diagnostics=diags,
                         diagnostics_every=50)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        E = np.array([d[2] + d[3] for d in diags]); P = np.array([d[4:7] for d in diags])
        p_scale = 1e9 * np.sqrt(2 * diags[0][2] / 1e9)
        out += [f"## C4: Plummer N={N}, float64, {kernel}, softening 0.02, {steps} KDK steps dt=1e-4 "
                f"(t_dyn = 0.0152 -> {steps*1e-4/0.0152:.1f} t_dyn)\n",
                f"wall {dt:.1f} s = {dt/ (steps + 1) * 1e3:.1f} ms per force evaluation incl. set-up "
                f"({N*N*(steps+1)/dt/1e9:.0f} Gint/s); virial ratio at t=0: {diags[0][2]/abs(diags[0][3]):.4f}\n",
                "| step | KE | PE | (E-E0)/|E0| | |p|/(M sigma) |", "|---:|---:|---:|---:|---:|"]
        for d, e, p in zip(diags, E, P):
            out.append(f"| {d[0]} | {d[2]:.6e} | {d[3]:.6e} | {(e - E[0]) / abs(E[0]):+.3e} | {np.linalg.norm(p)/p_scale:.2e} |")
        out.append(f"\nmax |dE/E0| = {np.abs(E - E[0]).max() / abs(E[0]):.3e}\n")

steps5 = arg("--c5-steps", 5)
if steps5 > 0:
    xd, md = nb.make_plummer_sphere(1_600_000, M_total=1e9, a=1.0, seed=1)
    xs, ms = nb.make_plummer_sphere(400_000, M_total=2e7, a=0.3, seed=2)
    xvm = np.vstack([xd, xs, np.zeros((1, 6))])
    sp = [nb.Species.dark(1_600_000, mass=float(md[0]), softening=0.05),
          nb.Species.star(400_000, mass=float(ms[0]), softening=0.02), nb.Species("bh", 1, 1e8, softening=0.001)]
    import warnings

    diags = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nb.run_simulation(xvm, sp, 0.0, 1e-4, 1e-4, precision="float32", save_snapshots=False, verbose=False,
                          output_dir=f"/tmp/c5_{rank}")          # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = nb.run_simulation(xvm, sp, 0.0, steps5 * 1e-4, 1e-4, precision="float32", save_snapshots=False,
                                verbose=False, output_dir=f"/tmp/c5_{rank}", diagnostics=diags,
                                diagnostics_every=steps5)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    Nn = len(xvm)
    E = [d[2] + d[3] for d in diags]
    out += [f"## C5: dark 1.6M + star 400k + BH (N={Nn}), float32, spline via run_simulation, {steps5} steps\n",
            f"wall {dt:.2f} s for {steps5 + 1} force+potential evaluations incl. set-up, H2D and the final D2H + restart "
            f"file: {dt/(steps5+1)*1e3:.0f} ms each ({Nn*Nn*(steps5+1)/dt/1e9:.0f} Gint/s aggregate); "
            f"species out: {[(k, v.shape) for k, v in res.items()]}; dE/E0 over the run: {(E[-1]-E[0])/abs(E[0]):+.2e}\n"]
if rank == 0:
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/config_runs_{world}gpu.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))
if world > 1:
    dist.destroy_process_group()
