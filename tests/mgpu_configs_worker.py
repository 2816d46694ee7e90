"""Worker for tests/test_gpu_multi.py::test_baseline_configs_sharded — run under torchrun, one rank per GPU.

BASELINE.json configurations 3, 4 and 5 at FULL size, sharded over all ranks of the job, checked against the
float64 oracle on row subsets (the reference kernels are independent per i, so rows against all N are exact;
SURVEY.md 8d):

  C3  Plummer N=1M, float32_kahan, force+potential, plummer and spline        rel-RMS <= 1e-5
  C4  Plummer N=262144, float64, plummer and spline                           max_i |da|/|a| <= 1e-12
  C5  dark 1.6M + star 400k + 1 BH (N=2,000,001), float32, spline              rel-RMS <= 1e-5

Both entry points are exercised: the fields API with group= (every rank gets all rows back) and the resident
DirectEngine that run_nbody_gpu steps (each rank holds the rows of its shard).  Rows are taken at the start,
across a shard boundary and at the end of the particle range.  Writes result.txt (OK/FAIL + one line per check).
"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nbody_streams_b200 as nb  # noqa: E402
import oracle  # noqa: E402
from nbody_streams_b200._lib import MODE_BOTH  # noqa: E402
from nbody_streams_b200.engine import DirectEngine, ShardLayout  # noqa: E402
from nbody_streams_b200.fields import _PRECISION_MAP, EPS2_DEFAULT, KERNEL_MAP  # noqa: E402

G = nb.G_DEFAULT


def rel_rms(a, r):
    a, r = np.asarray(a, np.float64), np.asarray(r, np.float64)
    return float(np.sqrt(((a - r) ** 2).sum() / (r ** 2).sum()))


def rel_max(a, r):
    a, r = np.asarray(a, np.float64), np.asarray(r, np.float64)
    if a.ndim == 2:
        return float((np.linalg.norm(a - r, axis=1) / np.linalg.norm(r, axis=1)).max())
    return float((np.abs(a - r) / np.abs(r)).max())


def row_blocks(N, world, rows):
    """Three blocks of `rows` rows: the first rows, a block straddling the boundary of the middle shards, the last."""
    L = ShardLayout(N, world)
    mid = min(max(L.first(max(1, world // 2)) - rows // 2, 0), N - rows)
    return [(0, rows), (mid, mid + rows), (N - rows, N)]


def check_config(name, pos, mass, h, precision, kernels, rows, tol, rank, world, msgs):
    ok = True
    N = pos.shape[0]
    f64 = precision == "float64"
    def rnd(a):   # what the kernel sees: inputs cast to the kernel dtype (scalars stay scalars)
        if np.isscalar(a):
            return float(a) if f64 else float(np.float32(a))
        return np.asarray(a, np.float64) if f64 else np.asarray(a, np.float32).astype(np.float64)

    p64, m64, h64 = rnd(pos), rnd(mass), rnd(h)
    prec, _ = _PRECISION_MAP[precision]
    blocks = row_blocks(N, world, rows)
    L = ShardLayout(N, world)
    for kernel in kernels:
        t0 = time.perf_counter()
        acc, phi = nb.compute_nbody_forces_potential_gpu(pos, mass, h, G, precision=precision, kernel=kernel,
                                                         group=dist.group.WORLD)
        t_api = time.perf_counter() - t0
        # the resident engine (what run_nbody_gpu steps): own rows only
        eng = DirectEngine(N, mass, h, prec, KERNEL_MAP[kernel], G, EPS2_DEFAULT[precision],
                           torch.device("cuda", torch.cuda.current_device()))
        xv = np.hstack([pos, np.zeros_like(pos)])
        eng.load_state(xv)
        eng.compute(MODE_BOTH, positions_changed=False)
        torch.cuda.synchronize()
        a_loc, p_loc = eng.acc.cpu().numpy(), eng.phi.cpu().numpy()
        sl = L.rows(rank)
        d_eng = max(rel_max(a_loc, acc[sl]) if len(a_loc) else 0.0, rel_max(p_loc, phi[sl]) if len(p_loc) else 0.0)
        sym = eng.symmetric
        del eng
        d_p2p = 0.0
        if "--p2p" in sys.argv:
            # exchange='p2p': workspace in symmetric memory, reduce + finish fused into one kernel over NVLink peer memory
            eng = DirectEngine(N, mass, h, prec, KERNEL_MAP[kernel], G, EPS2_DEFAULT[precision],
                               torch.device("cuda", torch.cuda.current_device()), exchange="p2p")
            eng.load_state(xv)
            eng.compute(MODE_BOTH, positions_changed=False)
            eng.compute(MODE_BOTH, positions_changed=False)      # twice: the workspace is cleared and reused
            torch.cuda.synchronize()
            a_p, p_p = eng.acc.cpu().numpy(), eng.phi.cpu().numpy()
            d_p2p = max(rel_max(a_p, acc[sl]) if len(a_p) else 0.0, rel_max(p_p, phi[sl]) if len(p_p) else 0.0)
            del eng
            d_eng = max(d_eng, d_p2p)
        if rank == 0:
            errs = []
            for (r0, r1) in blocks:
                kw = dict(kernel=kernel, rows=(r0, r1), fast=True)
                a_ref = oracle.compute_nbody_forces_cpu(p64, m64, h64, G, **kw)
                p_ref = oracle.compute_nbody_potential_cpu(p64, m64, h64, G, **kw)
                if f64:
                    errs.append(max(rel_max(acc[r0:r1], a_ref), rel_max(phi[r0:r1], p_ref)))
                else:
                    errs.append(max(rel_rms(acc[r0:r1], a_ref), rel_rms(phi[r0:r1], p_ref)))
            good = max(errs) <= tol and np.isfinite(acc).all() and np.isfinite(phi).all()
            msgs.append(f"{name} world={world} {precision} {kernel} ({'symmetric' if sym else 'one-sided'} sweep): "
                        f"{'rel-max' if f64 else 'rel-RMS'} vs oracle rows {[f'{e:.2e}' for e in errs]} (tol {tol:g}), "
                        f"engine (nccl{' and p2p' if '--p2p' in sys.argv else ''}) vs API rows {d_eng:.1e}, "
                        f"API call {t_api * 1e3:.0f} ms -> {'ok' if good else 'FAIL'}")
            ok &= bool(good)
        ok &= bool(d_eng <= (1e-12 if f64 else 5e-6))
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = Path(sys.argv[1])
    which = sys.argv[2].split(",") if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ["c3", "c4", "c5"]
    oracle.build()
    oracle.set_num_threads(len(os.sched_getaffinity(0)), fast=True)     # torchrun sets OMP_NUM_THREADS=1
    ok, msgs = True, []
    if "c3" in which:
        xv, m = nb.make_plummer_sphere(1_000_000, M_total=1e9, a=1.0, seed=42069)
        ok &= check_config("C3", xv[:, :3], float(m[0]), 0.01, "float32_kahan", ("plummer", "spline"), 2048, 1e-5,
                           rank, world, msgs)
    if "c4" in which:
        xv, m = nb.make_plummer_sphere(262_144, M_total=1e9, a=1.0, seed=42069)
        ok &= check_config("C4", xv[:, :3], float(m[0]), 0.02, "float64", ("spline", "plummer"), 1024, 1e-12,
                           rank, world, msgs)
    if "c5" in which:
        xd, md = nb.make_plummer_sphere(1_600_000, M_total=1e9, a=1.0, seed=1)
        xs, ms = nb.make_plummer_sphere(400_000, M_total=2e7, a=0.3, seed=2)
        pos = np.vstack([xd[:, :3], xs[:, :3], np.zeros((1, 3))])
        mass = np.concatenate([md, ms, [1e8]])
        h = np.concatenate([np.full(1_600_000, 0.05), np.full(400_000, 0.02), [0.001]])
        ok &= check_config("C5", pos, mass, h, "float32", ("spline",), 2048, 1e-5, rank, world, msgs)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.mkdir(parents=True, exist_ok=True)
        (out / "result.txt").write_text(("OK\n" if flag.item() == 1.0 else "FAIL\n") + "\n".join(msgs) + "\n")
        print("\n".join(msgs), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
