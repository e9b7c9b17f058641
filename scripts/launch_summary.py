"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share of the step (cold-cache, serialised: compare SHARES, not absolutes)."""
import collections
import csv
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    agg = collections.OrderedDict()
    n = 0
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        n += 1
        if n <= skip:
            continue
        unit = d["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v * 1e6 if unit == "s" else v
        name = d["Kernel Name"].split("(")[0][-48:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':50s} {'launches':>8s} {'total_us':>10s} {'share':>6s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:50s} {a[0]:8d} {a[1]:10.1f} {a[1] / tot:6.3f}")
    print(f"{'TOTAL':50s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
