"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def summarise(path, skip=()):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        k = row["Kernel Name"]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(row["Metric Unit"], 1)
        if any(s in k for s in skip):
            continue
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    return agg, tot


if __name__ == "__main__":
    skip = ("k_rmat", "cub::", "k_degm", "k_distinct", "k_scatter", "k_row_sectors", "k_labels", "k_build_sig", "k_csr")
    agg, tot = summarise(sys.argv[1], skip if len(sys.argv) > 2 else ())
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        print("%-90s n=%5d %10.3f ms %5.1f%%" % (k[:90], n, t / 1e6, 100 * t / tot))
    print("total %.3f ms" % (tot / 1e6))
