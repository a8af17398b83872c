#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: mean duration per (kernel, grid)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name).replace("<unnamed>::", "")
        key = (name[:58], row["Grid Size"], row["Block Size"])
        agg.setdefault(key, []).append(v)
    total = sum(sum(v) for v in agg.values())
    lines = ["%-58s %-14s %-12s %5s %9s %9s %6s" % ("kernel", "grid", "block", "n", "mean_us", "sum_us", "share")]
    for (name, grid, block), v in agg.items():
        lines.append("%-58s %-14s %-12s %5d %9.2f %9.1f %5.1f%%" % (name, grid.replace(" ", ""), block.replace(" ", ""), len(v),
                                                                  sum(v) / len(v), sum(v), 100 * sum(v) / total))
    lines.append("total %.1f us over %d launches" % (total, sum(len(v) for v in agg.values())))
    text = "\n".join(lines)
    print(text)
    if out:
        with open(out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
