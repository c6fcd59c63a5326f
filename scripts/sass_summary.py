#!/usr/bin/env python
"""Static evidence about the compiled kernels (no GPU needed): registers, shared memory and the memory / barrier /
TMA mnemonics of every kernel in libcudecomp.so, from cuobjdump. Written to profiles/ next to the measured numbers.
    python scripts/sass_summary.py > profiles/r1_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cudecomp_b200", "lib", "libcudecomp.so")
KEEP = re.compile(r"^(LDG|STG|LDS|STS|LD\.|ST\.|ATOM|RED|MEMBAR|FENCE|BAR|UBLKCP|UTMA|SYNCS|ERRBAR|CCTL|NANOSLEEP|LDC|S2R|S2UR)")


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)), local=int(m.group(5)))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, ops = None, collections.defaultdict(collections.Counter)
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            total[cur] += 1
            op = m.group(1)
            if KEEP.match(op):
                ops[cur][op] += 1
    print("# Static summary of the kernels in cudecomp_b200/lib/libcudecomp.so (sm_100a), from cuobjdump -res-usage / -sass.")
    print("# Not a measurement: what the compiler produced. LDG/STG.E.128 = 16-byte accesses, .ENL2.256 = 32-byte accesses,")
    print("# UBLKCP = cp.async.bulk (TMA unit), SYNCS = mbarrier operations, .EF = evict-first (streaming) cache hint.\n")
    for fn in sorted(usage, key=demangle):
        u = usage[fn]
        print("%s" % demangle(fn))
        print("    registers %d, static shared %d B, stack %d B, local %d B, %d SASS instructions" % (
            u["reg"], u["shared"], u["stack"], u["local"], total.get(fn, 0)))
        print("    " + ", ".join("%s x%d" % kv for kv in sorted(ops.get(fn, {}).items())))
        print()


if __name__ == "__main__":
    main()
