#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])[:100]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, t, t / n, 100 * t / tot))
    print("\ntotal %.1f us over %d launches (cold-cache, serialised: compare shares, not absolutes)" % (tot, len(data)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
