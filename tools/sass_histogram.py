#!/usr/bin/env python
"""Opcode histogram per kernel of libmtdgan_sm100a.so from `cuobjdump -sass`: the SASS evidence for tcgen05 / TMEM / TMA
(UTCHMMA / UTCQMMA, LDTM / STTM, UTMALDG / UTMASTG, UBLKCP) and warp shuffles (SHFL) per kernel.
Usage: python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "SHFL", "HMMA",
       "FFMA", "LDS", "STS", "LDG", "STG", "ATOMG", "REDG", "RED", "ATOMS", "BAR", "MUFU", "DFMA", "ACQBULK", "ELECT")


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(anonymous namespace\)::|^void\s+", "", kern).split("(")[0]
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            hist[kern][m.group(1)] += 1
    print(f"# cuobjdump -sass {os.path.basename(lib)}: instruction counts per kernel (selected opcodes; total = all instructions)")
    keys = [k for k in KEY if any(op.startswith(k) for h in hist.values() for op in h)]
    print(f"{'kernel':52s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in keys))
    for k, h in hist.items():
        print(f"{k[:52]:52s} {sum(h.values()):6d} " + " ".join(f"{sum(v for op, v in h.items() if op.startswith(key)):7d}" for key in keys))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mtd-gan_b200", "libmtdgan_sm100a.so"))
