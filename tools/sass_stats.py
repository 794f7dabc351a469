#!/usr/bin/env python
"""Static SASS statistics of the kernels in an object file (needs only `cuobjdump`; no GPU).

  python tools/sass_stats.py build/kernels_fast.o [substring ...]

Per kernel: instruction count, registers are in build/*.ptxas.log; here the opcode histogram and the longest backward-branch
loop body (a proxy for the instructions per iteration of the hot loop).  Used for profiles/r1_notes.md item 21, where the
instruction count of the inner loops predicted every measured win and loss."""
import collections
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    cur, d = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            d[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return d


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
    except Exception:
        return name


def stats(ins):
    ops = collections.Counter()
    loops = []
    for addr, text in ins:
        t = text.split()
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op.split(".")[0]] += 1
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:  # backward branch = loop
            lo = int(m.group(1), 16)
            loops.append(sum(1 for a, _ in ins if lo <= a <= addr))
    return ops, sorted(loops, reverse=True)


if __name__ == "__main__":
    obj, filters = sys.argv[1], sys.argv[2:]
    for name, ins in kernels(obj).items():
        dn = demangle(name)
        if filters and not any(f in dn or f in name for f in filters):
            continue
        ops, loops = stats(ins)
        top = ", ".join(f"{k} {v}" for k, v in ops.most_common(8))
        print(f"{dn[:100]}\n    {len(ins)} instructions; loop bodies (instructions): {loops[:4]}; {top}")
