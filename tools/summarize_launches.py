#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.   python tools/summarize_launches.py <csv>"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(name, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | min | max | share |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.2f | %.2f | %.1f %% |" % (k[:60], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, min(a[2]) / 1e3, max(a[2]) / 1e3, 100 * a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
