#!/usr/bin/env python3
"""Device time of each launch of one KDK step at small N (GPU box), and of the graph replay of the step.
`python tools/small_n_breakdown.py > gpurun_out/small_n_breakdown.md`"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200._lib import MODE_FORCE  # noqa: E402
from nbody_streams_b200.engine import DirectEngine  # noqa: E402
from nbody_streams_b200.fields import EPS2_DEFAULT, KERNEL_MAP, _PRECISION_MAP  # noqa: E402


def dev_time(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # microseconds


def main():
    print("| N | precision | kernel | kick+drift+pack us | sweep us | finish us | kick us | sum us | graph replay us |")
    print("|---:|---|---|---:|---:|---:|---:|---:|---:|")
    for N in (1_000, 4_000, 10_000, 40_000):
        xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=1)
        for prec_name, kernel in (("float32_kahan", "spline"), ("float32_kahan", "plummer"), ("float64", "spline")):
            prec = _PRECISION_MAP[prec_name][0]
            eng = DirectEngine(N, float(m[0]), 0.05, prec, KERNEL_MAP[kernel], nb.G_DEFAULT, EPS2_DEFAULT[prec_name],
                               torch.device("cuda", 0))
            eng.load_state(xv)
            eng.compute(MODE_FORCE)
            dt = 1e-7
            grid = eng._grid(MODE_FORCE)
            ws = eng._workspace(grid)
            n_tiles = eng.layout.tiles_per_rank
            import ctypes as C
            n_jt = (C.c_int * 4)(n_tiles, 0, 0, 0)
            reps = 300
            t_kd = dev_time(lambda: eng.kick_drift(dt), reps)
            t_sw = dev_time(lambda: eng._sweep(MODE_FORCE, 0, n_tiles, 0, 0, ws, 0, grid), reps)
            t_fi = dev_time(lambda: eng.lib.nb_finish(eng.prec, eng.kernel_id, MODE_FORCE, eng.n_local, grid, 1, n_jt,
                                                      ws.data_ptr(), eng.G, eng.acc.data_ptr(), None, eng._stream()), reps)
            t_k = dev_time(lambda: eng.kick(dt), reps)
            step = eng.capture_step(MODE_FORCE, dt)
            t_g = dev_time(step, 1000)
            print(f"| {N} | {prec_name} | {kernel} | {t_kd:.1f} | {t_sw:.1f} | {t_fi:.1f} | {t_k:.1f} | "
                  f"{t_kd + t_sw + t_fi + t_k:.1f} | {t_g:.1f} |", flush=True)


if __name__ == "__main__":
    main()
