#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel counts, total and share."""
import collections
import csv
import re
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    cols = rows[hdr]
    ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
    by = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        short = re.sub(r'\(anonymous namespace\)::|<unnamed>::|^void\s+', '', r[ki].split('(')[0])
        v = float(r[vi].replace(',', ''))
        us = v / 1000 if r[ui] in ('ns', 'nsecond') else (v if r[ui] in ('us', 'usecond') else v * 1000)
        by[short][0] += 1
        by[short][1] += us
        tot += us
    n = sum(c for c, _ in by.values())
    print(f"{n} launches, {tot / 1000:.3f} ms total (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':64s} {'count':>6s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
    for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[:64]:64s} {c:6d} {t / 1000:9.3f} {100 * t / tot:6.1f}% {t / c:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
