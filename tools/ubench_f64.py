#!/usr/bin/env python3
"""FP64-pipe microbenchmarks (csrc/nb_ubench.cu 7, 17-20): lane-ops/s and issue cycles per warp instruction per SM sub-partition."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from nbody_streams_b200 import _lib  # noqa: E402

lib = _lib.load()
torch.cuda.init()
sms = torch.cuda.get_device_properties(0).multi_processor_count
names = {7: "DFMA a=fma(a,b,c), b c loop-invariant", 17: "DFMA acc=fma(f,d,acc), f shared by a trio", 18: "DFMA acc=fma(d,d,acc)",
         19: "DFMA acc=fma(f,d,acc), all distinct", 20: "DADD acc=d-acc"}
print("| probe | lane-ops/s | TFLOP/s (x2) | cycles per warp instruction per sub-partition at 1965 MHz |")
print("|---|---:|---:|---:|")
for which, name in names.items():
    best = 0.0
    for _ in range(3):
        ops, ms = C.c_double(), C.c_double()
        _lib.check(lib.nb_ubench(which, 20000, C.byref(ops), C.byref(ms)))
        best = max(best, ops.value)
    cyc = 1.965e9 * sms * 4 * 32 / best
    print(f"| {which}: {name} | {best:.4e} | {2 * best / 1e12:.2f} | {cyc:.2f} |", flush=True)
