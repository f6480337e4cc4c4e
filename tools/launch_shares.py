"""Shares per kernel of an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file x.csv):
python tools/launch_shares.py x.csv > profiles/..._launches.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)       # -> microseconds
    tot[r[ki]] += v
    cnt[r[ki]] += 1
s = sum(tot.values())
print("# totals in microseconds")
for k, v in tot.most_common():
    print("%6.2f%%  launches=%4d  total=%12.1f  %s" % (100 * v / s, cnt[k], v, k[:180]))
