#!/usr/bin/env python3
"""A/B of experimental library builds (tools/build_variant.py) on one GPU: symmetric sweep times and the difference
to the one-sided sweep of the same library.

    python tools/sym_ab.py --libs base,d43u2,... [--prec float64] [--n 262144] [--out gpurun_out/sym_ab.md]

Each library runs in its own process (NBODY_SM100_LIB=tmp_ab/lib_<name>.so; 'intree' = the in-tree build)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def child(prec_key, N):
    import numpy as np
    import torch

    import nbody_streams_b200 as nb
    from nbody_streams_b200.fields import _PRECISION_MAP, _run_direct

    prec, dtype_np = _PRECISION_MAP[prec_key]
    xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
    dt = np.float64 if prec_key == "float64" else np.float32
    pos_t = torch.from_numpy(np.ascontiguousarray(xv[:, :3].astype(dt))).cuda()
    eps2 = 1e-30 if prec_key == "float64" else 1e-15
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    cells = []
    for kernel in ("plummer", "spline"):
        kid = nb.KERNEL_MAP[kernel]
        for mode in (1, 3):
            def run(alg):
                return _run_direct(pos_t, None, float(m[0]), None, 0.02, N, prec, kid, mode, nb.G_DEFAULT, eps2, dtype_np,
                                   algorithm=alg)
            ms = []
            for r in range(4):
                ev[0].record()
                out = run("symmetric")
                ev[1].record()
                torch.cuda.synchronize()
                if r:
                    ms.append(ev[0].elapsed_time(ev[1]))
            ref = run("onesided")
            a, b = out[0].double(), ref[0].double()
            err = float(((a - b).norm(dim=1) / b.norm(dim=1)).max())
            cells.append(f"{np.median(ms):.2f} ({err:.0e})")
    print("ROW " + " | ".join(cells), flush=True)


def main():
    if "--child" in sys.argv:
        return child(arg("--prec", "float64"), arg("--n", 262144))
    libs = arg("--libs", "intree").split(",")
    out = Path(arg("--out", "gpurun_out/sym_ab.md"))
    lines = [f"symmetric sweep ms (max rel diff of acc to the one-sided sweep), {arg('--prec', 'float64')} N = {arg('--n', 262144)}",
             "", "| build | plummer force | plummer both | spline force | spline both |", "|---|---|---|---|---|"]
    for name in libs:
        env = dict(os.environ)
        if name != "intree":
            env["NBODY_SM100_LIB"] = str(ROOT / "tmp_ab" / f"lib_{name}.so")
        r = subprocess.run([sys.executable, __file__, "--child", "--prec", arg("--prec", "float64"), "--n", str(arg("--n", 262144))],
                           env=env, capture_output=True, text=True, timeout=600)
        row = [l[4:] for l in r.stdout.splitlines() if l.startswith("ROW ")]
        lines.append(f"| {name} | {row[0] if row else 'FAILED: ' + r.stderr[-300:].replace(chr(10), ' ')} |")
        print(lines[-1], flush=True)
    out.parent.mkdir(exist_ok=True)
    out.write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
