"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches / total us / share / avg us
(markdown table on stdout).  Developer tool used for profiles/*_launches_summary.md."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:80]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.2f} |")


if __name__ == "__main__":
    main(sys.argv[1])
