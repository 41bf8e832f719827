#!/usr/bin/env python
"""Static instruction counts of the innermost loops of a kernel (cuobjdump -sass), to compare kernel formulations
without a GPU: prints registers/stack of every kernel in an object file and, per kernel, every innermost loop that
contains a given mnemonic (default SHFL.UP = the 32-column chunk loops of the row sweeps in K4/K5).

    python tools/sass_loops.py build/cuda/bwag_extend.o [MNEMONIC]
"""
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-res-usage", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    res, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            res[name] = tuple(int(x) for x in m.groups())
    return res


def loops(obj, fn, needle):
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], stdout=subprocess.PIPE, text=True).stdout
    ins = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = [a for a, _ in ins]
    back = []
    for a, text in ins:
        m = re.search(r"\bBRA\b(?:\.\w+)*\s+(?:[!\w]+,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < a:
            back.append((int(m.group(1), 16), a))
    found = []
    for t, a in back:
        if any(t < t2 and a2 < a for t2, a2 in back):   # not innermost
            continue
        body = [x for ad, x in ins if t <= ad <= a]
        if any(needle in x for x in body):
            found.append((t, a, len(body)))
    return found, len(addr)


def main():
    obj = sys.argv[1]
    needle = sys.argv[2] if len(sys.argv) > 2 else "SHFL.UP"
    for fn, (reg, stack, shared) in sorted(kernels(obj).items()):
        fl, n = loops(obj, fn, needle)
        print("%s: %d instructions, %d registers, %d B stack, %d B static shared" % (fn, n, reg, stack, shared))
        for t, a, k in fl:
            print("    innermost loop with %s at 0x%04x-0x%04x: %d instructions" % (needle, t, a, k))


if __name__ == "__main__":
    main()
