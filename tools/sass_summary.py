#!/usr/bin/env python
"""Per-kernel SASS instruction counts of libscanerf_b200.so (cuobjdump -sass): which kernels carry the tensor-core
(UTCHMMA / UTCBAR / LDTM = tcgen05.mma / commit / ld), reduction (RED / REDG / ATOM) and conversion instructions.
  python tools/sass_summary.py [profiles/r2_sass_summary.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = [d for d in os.listdir(ROOT) if d.endswith("_b200")][0]
LIB = os.path.join(ROOT, PKG, "lib", "libscanerf_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "RED", "ATOM", "MUFU", "F2FP", "SHFL", "LDG", "STG", "LDS", "STS", "BAR", "SYNCS"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.md")
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    counts, total, cur = collections.OrderedDict(), collections.Counter(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)[:70]
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            total[cur] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
                    break
    with open(out, "w") as f:
        f.write(f"# SASS instruction counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)\n\n")
        f.write("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld (TMEM -> registers), RED/ATOM = global / shared reductions.\n\n")
        f.write("| kernel | instructions | " + " | ".join(KEYS) + " |\n|---|---|" + "---|" * len(KEYS) + "\n")
        for k, c in counts.items():
            if total[k] == 0:
                continue
            f.write(f"| `{k}` | {total[k]} | " + " | ".join(str(c.get(x, 0)) for x in KEYS) + " |\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
