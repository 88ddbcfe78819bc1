"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    val = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    val = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
    agg[name][0] += 1
    agg[name][1] += val
    tot += val
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d total=%9.1f us  avg=%8.1f us  share=%5.1f%%" % (k[:62], n, t, t / n, 100 * t / tot))
print("total %.1f us" % tot)
