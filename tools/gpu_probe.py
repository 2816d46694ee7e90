#!/usr/bin/env python3
"""GPU probe: peak microbenchmarks, tuning-variant timings and parity numbers in one call.

    python tools/gpu_probe.py [--n 262144] [--out gpurun_out/probe.json]

Writes one JSON document; nothing here is a bench value (see bench.py)."""
from __future__ import annotations

import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200 import _lib  # noqa: E402
from nbody_streams_b200.fields import _run_direct, _PRECISION_MAP  # noqa: E402


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def time_direct(pos_t, m, h, N, prec_key, kernel, mode, reps=3):
    prec, dtype_np = _PRECISION_MAP[prec_key]
    kid = nb.KERNEL_MAP[kernel]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for r in range(reps + 1):
        ev[0].record()
        _run_direct(pos_t, None, m, None, h, N, prec, kid, mode, nb.G_DEFAULT, 1e-15 if prec != 2 else 1e-30, dtype_np)
        ev[1].record()
        torch.cuda.synchronize()
        if r:
            ms.append(ev[0].elapsed_time(ev[1]))
    return float(np.median(ms))


def main():
    out = {"device": nb.get_gpu_info()}
    lib = _lib.load()
    N = arg("--n", 262144)
    # ---- peaks
    names = {0: "ffma", 1: "ffma2", 2: "ffma2_3op", 3: "fadd2", 4: "fmul2", 5: "mufu_rsq", 6: "mix_packed",
             7: "dfma", 8: "mix_scalar", 9: "ffma2x12+mufu2", 10: "ffma2x12+fmnmx4", 11: "ffmax12+mufu2",
             12: "ffma2x12+mufu2+fmnmx2", 13: "loop_both_uni", 14: "loop_both_uni_dadd", 15: "loop_force_uni",
             16: "ffma2_trio_shared_f"}
    ub = {}
    for which, name in names.items():
        ops, ms = C.c_double(), C.c_double()
        _lib.check(lib.nb_ubench(which, 40000 if which in (6, 8, 13, 14, 15) else 20000, C.byref(ops), C.byref(ms)))
        ub[name] = {"per_s": ops.value, "ms": ms.value}
        print(name, f"{ops.value:.4e}/s", f"{ms.value:.2f} ms", flush=True)
    out["ubench"] = ub
    # ---- variants
    xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
    pos_t = torch.from_numpy(xv[:, :3].astype(np.float32)).cuda()
    pos64 = torch.from_numpy(np.ascontiguousarray(xv[:, :3])).cuda()
    res = {}
    for v in ([int(x) for x in arg('--variants', '0,1,2,3').split(',')]):
        lib.nb_set_variant(0, v)
        for kernel in ("plummer", "spline"):
            for prec in ("float32", "float32_kahan"):
                for mode, mname in ((1, "force"), (3, "both"), (2, "pot")):
                    try:
                        ms = time_direct(pos_t, float(m[0]), 0.01, N, prec, kernel, mode)
                    except Exception as e:  # noqa: BLE001
                        res[f"v{v}_{kernel}_{prec}_{mname}"] = str(e)
                        print("ERR", v, kernel, prec, mname, e, flush=True)
                        continue
                    gi = N * N / ms / 1e6
                    res[f"v{v}_{kernel}_{prec}_{mname}"] = {"ms": ms, "gint_s": gi}
                    print(f"v{v} {kernel:8s} {prec:14s} {mname:5s} {ms:9.3f} ms  {gi:9.1f} Gint/s", flush=True)
    lib.nb_set_variant(0, -1)
    for v in range(3):
        lib.nb_set_variant(2, v)
        for kernel in ("plummer", "spline"):
            for mode, mname in ((1, "force"), (3, "both")):
                ms = time_direct(pos64, float(m[0]), 0.01, N, "float64", kernel, mode, reps=2)
                gi = N * N / ms / 1e6
                res[f"f64_v{v}_{kernel}_{mname}"] = {"ms": ms, "gint_s": gi}
                print(f"f64 v{v} {kernel:8s} {mname:5s} {ms:9.3f} ms  {gi:9.1f} Gint/s", flush=True)
    lib.nb_set_variant(2, -1)
    for kernel in ("newtonian", "dehnen_k1", "dehnen_k2"):
        ms = time_direct(pos_t, float(m[0]), 0.01, N, "float32", kernel, 1)
        res[f"v0_{kernel}_float32_force"] = {"ms": ms, "gint_s": N * N / ms / 1e6}
        print(f"v0 {kernel} float32 force {ms:9.3f} ms {N*N/ms/1e6:9.1f} Gint/s", flush=True)
    out["timings"] = {"N": N, **res}
    Path(arg("--out", "gpurun_out/probe.json")).parent.mkdir(parents=True, exist_ok=True)
    Path(arg("--out", "gpurun_out/probe.json")).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
