"""Summarise an ncu launch list (gpu__time_duration.sum csv) by kernel:
python tools/launch_summary.py launches.csv [last_n_launches]"""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
if len(sys.argv) > 2:
    n = int(sys.argv[2]); skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = rows[len(rows) - n - skip: len(rows) - skip]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    t = tot[name]
    t[0] += 1
    t[1] += float(r["Metric Value"].replace(",", ""))
s = sum(v[1] for v in tot.values())
print("launches %d total %.3f ms" % (len(rows), s / 1e6))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-48s n=%4d total=%9.3f ms avg=%9.1f us share=%5.1f%%" % (k[:48], v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / s))
