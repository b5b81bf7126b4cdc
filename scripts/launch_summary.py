"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        agg.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) / len(v) for v in agg.values())
    print(f"{'kernel':62s} {'n':>3s} {'avg us':>9s} {'share':>6s}")
    for k, v in agg.items():
        print(f"{k[:62]:62s} {len(v):3d} {sum(v) / len(v):9.1f} {sum(v) / len(v) / tot * 100:5.1f}%")
    print(f"sum of per-kernel averages: {tot:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
