"""Code-generation invariants of the shipped kernels, checked in the SASS of the built library (no GPU needed).

The hot loops are written for specific Blackwell instructions; a compiler flag, a header change or a toolkit
update that silently loses them would not fail any parity test, only the bench.  These tests pin what DESIGN.md
3.0 / 3.1 claim about the machine code (profiles/r02_sass_hist_kernels.txt is the tracked listing):
packed FP32 arithmetic (FFMA2 / FADD2 / FMUL2), the j-tile ring fed by TMA bulk copies (UBLKCP + mbarrier SYNCS),
double atomics for the pair-symmetric partial sums (REDG.E.ADD.F64), no local-memory traffic in the uniform
hot loops, and the per-step instruction budget of the bench kernel.
"""
import re
import shutil
import sys
from collections import Counter
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
BUILD = ROOT / "nbody_streams_b200" / "build"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not (BUILD / "nb_sym_k1.o").exists(),
                                reason="needs cuobjdump and the object files of an in-tree build")


def _kernels(obj: str):
    import sass_rf_model as M

    return {name: rows for name, rows in M.kernels(str(BUILD / obj))}


def _innermost_loops(rows):
    import sass_rf_model as M

    loops = []
    for addr, ins in rows:
        m = re.search(r"0x([0-9a-f]+)\s*$", ins) if "BRA" in ins else None
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    out = []
    for tgt, addr in loops:
        if any(t2 >= tgt and a2 <= addr and (t2, a2) != (tgt, addr) for t2, a2 in loops):
            continue
        body = [(a, i) for a, i in rows if tgt <= a <= addr]
        out.append((body, Counter(M.opcode(i) for _, i in body)))
    return out


@pytest.fixture(scope="module")
def sym_k1():
    return _kernels("nb_sym_k1.o")


def _find(kernels, pattern):
    hits = [n for n in kernels if re.search(pattern, n)]
    assert len(hits) == 1, (pattern, hits)
    return kernels[hits[0]]


def test_bench_kernel_hot_loop_is_packed_fp32_within_its_instruction_budget(sym_k1):
    """nb::sym_f32_kernel<plummer, force+potential, NP=4, 4 CTAs/SM, whole tiles>: the uniform fast path."""
    import sass_rf_model as M

    rows = _find(sym_k1, r"sym_f32_kernelILi1ELi3ELi4ELi4ELb0")
    loops = [(b, h) for b, h in _innermost_loops(rows) if h["SHFL"] >= 32 and h["FMNMX"] == 0 and h["FSEL"] == 0]
    assert len(loops) == 1, [dict(h) for _, h in loops]
    body, h = loops[0]
    steps = h["MUFU"] // 8                           # 8 reciprocal square roots per rotation step (4 packed pairs)
    assert steps == 8
    # per step: 36 FFMA2 + 20 FADD2 + 8 FMUL2 = 64 packed instructions, 8 SHFL (4 packed j-accumulators), 3 LDS
    assert (h["FFMA2"], h["FADD2"], h["FMUL2"]) == (36 * steps, 20 * steps, 8 * steps)
    assert h["SHFL"] == 8 * steps and h["LDS"] == 3 * steps
    assert h["LDL"] == 0 and h["STL"] == 0           # no spills inside the loop
    assert h["FFMA"] == 0 and h["FMUL"] == 0         # nothing fell back to scalar arithmetic
    other = len(body) - 64 * steps
    assert other <= 232, other                       # 224 in the final build of round 2 (247 before the offset rewrite)
    n, three, cycles = M.model(body)
    assert n == 64 * steps and cycles + other <= 1345, (three, cycles, other)   # 1330 modelled; 1375 measured


def test_symmetric_kernels_use_tma_and_double_atomics(sym_k1):
    for pat in (r"sym_f32_kernelILi1ELi3ELi4ELi4ELb0", r"sym_f64_kernelILi1ELi1ELi4ELi4ELb0"):
        ops = Counter(i.split()[0] if not i.startswith("@") else i.split()[1] for _, i in _find(sym_k1, pat))
        assert any(o.startswith("UBLKCP") for o in ops), pat            # cp.async.bulk (TMA engine) fills the ring
        assert any(o.startswith("SYNCS") for o in ops), pat             # mbarrier arrive / try_wait
        assert any(re.match(r"REDG?\.E\.ADD\.F64", o) for o in ops), pat   # double atomics into the workspace


def test_float64_force_loop_instruction_count(sym_k1):
    """18 FP64 instructions per pair in force-only sweeps: 3 DADD + 3 DFMA (r^2), 6 for r^-3 from the MUFU.RSQ64H
    seed, 6 accumulates; 4 i-particles per thread, 4 rotation steps per iteration."""
    rows = _find(sym_k1, r"sym_f64_kernelILi1ELi1ELi4ELi4ELb0")
    loops = [h for _, h in _innermost_loops(rows) if h["SHFL"] >= 12 and h["DSETP"] == 0]
    assert len(loops) == 1
    h = loops[0]
    pairs = h["MUFU"]
    assert pairs == 16 and h["DFMA"] + h["DADD"] + h["DMUL"] == 18 * pairs
    assert h["LDL"] == 0 and h["STL"] == 0


def test_one_sided_bench_kernel_hot_loop():
    """nb::sweep_f32_kernel (round-1 sweep, small N / n_sources): 2 j x 8 i per iteration, 96 packed instructions."""
    ks = _kernels("nb_sweep_k1.o")
    rows = _find(ks, r"sweep_f32_kernelILi1ELi3ELi1ELi2ELi5ELi4ELb0ELb0")
    loops = [h for _, h in _innermost_loops(rows) if h["FFMA2"] >= 48 and h["FMNMX"] == 0 and h["FSEL"] == 0]
    assert loops, "uniform hot loop not found"
    h = loops[0]
    assert (h["FFMA2"], h["FADD2"], h["FMUL2"], h["MUFU"]) == (48, 32, 16, 16)
    assert h["LDL"] == 0 and h["STL"] == 0
