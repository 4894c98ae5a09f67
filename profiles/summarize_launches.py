"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit in ('nsecond', 'ns') else (v * 1e3 if unit in ('msecond', 'ms') else v)
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('cvflow::', '')
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("kernels %d  launches %d  total %.1f us" % (len(agg), sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-56s n=%4d total=%9.1f us avg=%7.2f us share=%5.1f%%" % (k[:56], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
