#!/usr/bin/env python3
"""Build an experimental variant of libnbody_sm100.so into tmp_ab/ (git-ignored) without touching the in-tree library:

    python tools/build_variant.py NAME -DNB_SYM_NI_F64=2 -DNB_SYM_MINB_F64=6 ...

The translation units that see the symmetric-sweep build switches (nb_sym_k*.cu, nb_capi.cu) are recompiled with the
given flags; everything else is linked from nbody_streams_b200/build/.  Select it with NBODY_SM100_LIB=tmp_ab/lib_NAME.so.
Prints registers / spills of the sym kernels."""
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from nbody_streams_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    want = "sym_f64_kernel"
    if "--show" in flags:                       # --show SUBSTRING: which kernels' register counts to print
        k = flags.index("--show")
        want = flags[k + 1]
        del flags[k:k + 2]
    B.build()                                   # up-to-date objects of the unchanged units
    out = ROOT / "tmp_ab" / f"obj_{name}"
    out.mkdir(parents=True, exist_ok=True)
    redo = [s for s in B._sources() if s.stem.startswith("nb_sym_k") or s.stem == "nb_capi"]
    keep = [B.OBJ / (s.stem + ".o") for s in B._sources() if s not in redo]

    def cc(s):
        o = out / (s.stem + ".o")
        r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", str(s), "-o", str(o)], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr[-3000:])
        (out / (s.stem + ".ptxas.log")).write_text(r.stderr)
        return o

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(cc, redo))
    lib = ROOT / "tmp_ab" / f"lib_{name}.so"
    subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib),
                    *map(str, keep + objs), "-lcudart"], check=True)
    log = (out / "nb_sym_k1.ptxas.log").read_text() + (out / "nb_sym_k4.ptxas.log").read_text()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log):
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        if want in fn:
            print(f"{fn[:70]:70s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill st/ld {m.group(3)}/{m.group(4)}")
    print(lib)


if __name__ == "__main__":
    main()
