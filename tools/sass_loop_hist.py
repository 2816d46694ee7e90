#!/usr/bin/env python3
"""Histogram the opcodes of the hottest loop of each kernel in a cubin / .so.

Usage: sass_loop_hist.py <file> [name-substring] [--key OPC] [--all OPC]
Finds backward branches in `cuobjdump -sass` output, takes the loop body
[target, branch] with the most `--key` opcodes (default FFMA2, falls back to
DFMA/FFMA) and prints the opcode histogram, so the per-interaction issue-slot
budget can be checked without a GPU.  `--all OPC` prints EVERY innermost loop
that contains OPC instead (kernels with several tile paths: the uniform fast
path, the general per-particle path, the exact spline step).
"""
import re
import subprocess
import sys
from collections import Counter


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, rows = None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name:
                yield name, rows
            name, rows = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, rows


def opcode(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0].split(".")[0]


def main():
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    key = None
    if "--key" in sys.argv:
        key = sys.argv[sys.argv.index("--key") + 1]
    every = sys.argv[sys.argv.index("--all") + 1] if "--all" in sys.argv else None
    for name, rows in kernels(path):
        if sub not in name:
            continue
        if every:
            print(f"== {name}")
            loops = []
            for addr, ins in rows:
                m = re.search(r"0x([0-9a-f]+)\s*$", ins) if "BRA" in ins else None
                if m and int(m.group(1), 16) < addr:
                    loops.append((int(m.group(1), 16), addr))
            for tgt, addr in loops:
                if any(t2 >= tgt and a2 <= addr and (t2, a2) != (tgt, addr) for t2, a2 in loops):
                    continue                                  # not innermost
                c = Counter(opcode(i) for a, i in rows if tgt <= a <= addr)
                if c[every]:
                    print(f"   loop 0x{tgt:x}..0x{addr:x}: {sum(c.values())} instr")
                    print("      " + "  ".join(f"{o}:{n}" for o, n in c.most_common()))
            continue
        best = None
        for addr, ins in rows:
            if "BRA" not in ins:
                continue
            m = re.search(r"0x([0-9a-f]+)\s*$", ins)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt >= addr:
                continue
            body = [opcode(i) for a, i in rows if tgt <= a <= addr]
            c = Counter(body)
            k = key or ("FFMA2" if c["FFMA2"] else ("DFMA" if c["DFMA"] else "FFMA"))
            score = c[k] / max(1, sum(c.values())) if c[k] >= 8 else 0.0
            if best is None or score > best[0]:
                best = (score, tgt, addr, c, k)
        print(f"== {name}")
        if best is None:
            print("   (no loop)")
            continue
        score, tgt, addr, c, k = best
        total = sum(c.values())
        print(f"   loop 0x{tgt:x}..0x{addr:x}: {total} instr, {k}={c[k]}")
        print("   " + "  ".join(f"{o}:{n}" for o, n in c.most_common()))


if __name__ == "__main__":
    main()
