"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
        continue
    u = row.get('Metric Unit', '')
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1e-6)
    name = re.sub(r'<.*', '', row.get('Kernel Name', '').split('(')[0]).replace('void ', '').replace('pu::', '')
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f'# {sys.argv[1]}: {sum(c for c, _ in agg.values())} launches, {tot:.2f} ms total kernel time (cold-cache, serialised)')
print('ms\tshare\tlaunches\tkernel')
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{ms:.3f}\t{100 * ms / tot:.1f}%\t{c}\t{k}')
