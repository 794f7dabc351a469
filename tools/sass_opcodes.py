#!/usr/bin/env python
"""Per-kernel SASS opcode evidence of the in-tree library objects (needs only `cuobjdump`; no GPU):
   python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMA loads = UTMALDG, TMA stores = UTMASTG, bulk copies = UBLKCP, packed fp32x2 FMA = FFMA2."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "FFMA2"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o.replace("(anonymous namespace)::", "")).replace("void ", "") for o in out]


def main():
    print("# SASS opcode evidence per kernel of the in-tree library objects (cuobjdump -sass build/*.o; sm_100a); tools/sass_opcodes.py")
    print(f"{'kernel':92s} {'instr':>6s} " + " ".join(f"{c:>8s}" for c in COLS))
    tot = collections.Counter()
    for obj in sorted(glob.glob(os.path.join(ROOT, "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
        cur, d = None, collections.OrderedDict()
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                d[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur:
                d[cur][m.group(1)] += 1
                d[cur]["_n"] += 1
        names = demangle(list(d.keys()))
        tag = os.path.basename(obj)[:-2]
        for nm, c in zip(names, d.values()):
            if not any(c[k] for k in COLS) and c["_n"] < 400:
                continue  # small pointwise kernels without any of the listed opcodes
            print(f"{(tag + ': ' + nm)[:92]:92s} {c['_n']:6d} " + " ".join(f"{c[k]:8d}" for k in COLS))
            for k in COLS:
                tot[k] += c[k]
    print("# totals: " + ", ".join(f"{tot[k]} {k}" for k in COLS))


if __name__ == "__main__":
    main()
