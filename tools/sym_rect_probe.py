#!/usr/bin/env python3
"""Time single nb_sym_sweep launches with arbitrary tile ranges on ONE GPU (emulates what each rank of a
multi-GPU step launches): python tools/sym_rect_probe.py"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200 import _lib  # noqa: E402
from nbody_streams_b200.engine import ShardLayout, sym_plan  # noqa: E402

lib = _lib.load()
N = 1_000_000
prec, kid, mode = 1, 1, 3
xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
pos = torch.from_numpy(np.ascontiguousarray(xv[:, :3].astype(np.float32))).cuda()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L = ShardLayout(N, world)
n_slots = L.n_slots
packed = torch.zeros(lib.nb_packed_bytes(prec, n_slots), dtype=torch.uint8, device="cuda")
_lib.check(lib.nb_pack(prec, kid, packed.data_ptr(), 0, pos.data_ptr(), 0, None, 0, float(m[0]), None, 0, 0.01, N, n_slots,
                       1e-15, None))
ws = torch.zeros((n_slots, 4), dtype=torch.float64, device="cuda")
g = C.c_int(0)
_lib.check(lib.nb_sym_grid(prec, kid, mode, C.byref(g)))
grid = g.value
tpb = lib.nb_sym_block_tiles(prec)


def run(own, tri, A, B, Cr, ibC, reps=3, grid=grid):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for r in range(reps + 1):
        e0.record()
        _lib.check(lib.nb_sym_sweep(prec, kid, mode, packed.data_ptr(), N, own[0], own[1], tri, A[0], A[1], B[0], B[1],
                                    Cr[0], Cr[1], ibC[0], ibC[1], 1e-15, ws.data_ptr(), ws.numel() * 8, n_slots, 1, grid,
                                    None))
        e1.record()
        torch.cuda.synchronize()
        if r:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for rank in range(world):
    p = sym_plan(L, rank, tpb)
    t_tri = run(p["own"], 1, (0, 0), (0, 0), (0, 0), (0, 0))
    t_rect = run(p["own"], 0, p["A"], p["B"], p["C"], p["ibC"])
    t_all = run(p["own"], 1, p["A"], p["B"], p["C"], p["ibC"])
    print(f"world {world} rank {rank}: plan {p}  tri {t_tri:.2f} ms  rect {t_rect:.2f} ms  all-in-one {t_all:.2f} ms", flush=True)

if "--shapes" in sys.argv:
    # which property of the launch makes rank 1's rectangle slower?  same pair count, different shapes / grids
    for own, Cr, ibC in (((1954, 3907), (976, 1954), (0, 489)), ((1954, 3907), (0, 978), (0, 489)),
                         ((0, 1954), (2929, 3907), (0, 489)), ((0, 976), (1954, 3907), (0, 244)),
                         ((0, 1952), (1954, 2930), (0, 488)), ((0, 1952), (1954, 2954), (0, 488))):
        for gsz in (grid, grid - 4, 148 * 3):
            t = run(own, 0, (0, 0), (0, 0), Cr, ibC, grid=gsz)
            n_ib = min(ibC[1], (own[1] - own[0] + tpb - 1) // tpb)
            units = n_ib * (Cr[1] - Cr[0])
            print(f"own {own} C {Cr} ibC {ibC} grid {gsz}: {t:.2f} ms  {t * 1e3 * gsz / units:.1f} us/unit/cta", flush=True)
