"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and grid)."""
import collections
import csv
import re
import sys


def main(path, by_grid=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        if len(row) <= vi:
            continue
        try:
            v = float(row[vi].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", row[ki]).replace("void ", "").replace("<unnamed>::", "")[:48]
        if by_grid:
            k += " grid=" + row[gi]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if t / tot < 0.0005:
            continue
        print(f"| `{k}` | {n} | {t / 1e6:.3f} | {t / n / 1e3:.2f} | {t / tot * 100:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
