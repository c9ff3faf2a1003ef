"""Per-kernel durations of the last step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
idx = {h: i for i, h in enumerate(rows[hi])}
names = [(r[idx['Kernel Name']], float(r[idx['Metric Value']].replace(',', '')), r[idx['Metric Unit']])
         for r in rows[hi + 1:] if len(r) > idx['Metric Value'] and r[idx['Metric Name']] == 'gpu__time_duration.sum']
last = max(i for i, n in enumerate(names) if 'build_select' in n[0])
tot = 0.0
for n, v, u in names[last:]:
    us = v / 1000.0 if u in ('ns', 'nsecond') else v
    tot += us
    print('%-70s %8.1f us' % (n[:70], us))
print('sum %.1f us' % tot)
