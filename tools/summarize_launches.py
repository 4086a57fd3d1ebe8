"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel count, total time and share of the step."""
import csv
import collections
import re
import sys

path = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else '0'      # launches to skip (warm-up step); 'half' | 'after:<kernel name>'
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    v = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3) * v
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    rows.append((name, v))
if skip == 'half':
    rows = rows[len(rows) // 2:]
elif skip.startswith('after:'):          # everything after the first launch whose name contains the text
    key = skip[6:]
    first = next(i for i, (n, _) in enumerate(rows) if key in n)
    rows = rows[first + 1:]
else:
    rows = rows[int(skip):]
tot = sum(v for _, v in rows)
agg = collections.OrderedDict()
for n, v in rows:
    c, t = agg.get(n, (0, 0.0))
    agg[n] = (c + 1, t + v)
print('launches %d, total %.1f us' % (len(rows), tot))
print('%-60s %6s %12s %7s' % ('kernel', 'count', 'total_us', 'share'))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-60s %6d %12.1f %6.1f%%' % (n[:60], c, t, 100 * t / tot))
