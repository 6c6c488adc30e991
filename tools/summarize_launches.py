"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys

UNIT = {'nsecond': 1e-6, 'ns': 1e-6, 'usecond': 1e-3, 'us': 1e-3, 'msecond': 1, 'ms': 1, 'second': 1e3, 's': 1e3}
lines = [ln for ln in open(sys.argv[1]) if not ln.startswith('==')]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    tot[name] += float(row['Metric Value'].replace(',', '')) * UNIT[row['Metric Unit']]
    cnt[name] += 1
T = sum(tot.values())
print(f'{sum(cnt.values())} launches, {T:.1f} ms kernel time')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'{v:10.2f} ms {100 * v / T:6.2f}%  n={cnt[k]:5d}  {k}')
