"""Top stall locations of one kernel from `ncu --page source --csv --print-source cuda,sass` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]
idx = {}
for i, h in enumerate(hdr):
    idx.setdefault(h, i)


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(num(r[idx['# Samples']]) for r in data)
print('total samples', tot, 'rows', len(data))
for r in sorted(data, key=lambda r: -num(r[idx['# Samples']]))[:top]:
    print('%6d lsb=%-6s exec=%-9s thr=%-5s | %s | %s | %s' % (
        num(r[idx['# Samples']]), r[idx['stall_long_sb']], r[idx['Instructions Executed']],
        r[idx['Avg. Threads Executed']][:5], r[0], r[3][:60], r[1][:70]))
