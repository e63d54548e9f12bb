"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}.get(unit, 1e-6)
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'<.*', '', name)
    rows.append((name, v * scale, r['Kernel Name'], r.get('Grid Size', '')))
tot = sum(t for _, t, _, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, t, _, _ in rows:
    agg[n][0] += 1
    agg[n][1] += t
print(f'{len(rows)} launches, {tot:.3f} ms total (serialised, cold cache)')
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f'{t:9.3f} ms {100 * t / tot:5.1f}%  x{c:4d}  {n}')
print('--- top individual launches')
for n, t, full, grid in sorted(rows, key=lambda r: -r[1])[:25]:
    print(f'{t:9.3f} ms  {grid:>16s}  {full[:110]}')
