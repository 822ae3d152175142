"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_list.py file.csv [--seq]   (--seq prints every launch in order)"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    out = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")[:60]
        out.append((int(r[0]), n, r[gi], v / 1000.0))
    return out


if __name__ == "__main__":
    L = load(sys.argv[1])
    if "--seq" in sys.argv:
        for i, n, g, t in L:
            print(i, n, g, round(t, 1))
        sys.exit(0)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, n, _, t in L:
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.1f} us over {len(L)} launches")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:10.1f} us {100 * t / tot:5.1f}% {c:5d}  {n}")
