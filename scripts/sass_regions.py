"""Summarise an `ncu --page source --csv` dump: hot SASS regions by executed instructions and stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ie, si, ss = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
data = []
for r in rows[hi + 1:]:
    if len(r) > ie and r[ie].isdigit():
        data.append((r[si].strip(), int(r[ie]), int(r[ss])))
tot = sum(d[1] for d in data)
tots = sum(d[2] for d in data)
print("total inst", tot, "samples", tots, "n sass", len(data))
cur = None
start = 0
acc = accs = 0
segs = []
for idx, (s, c, sm) in enumerate(data):
    if cur is None:
        cur, start = c, idx
    if not (0.7 * cur <= c <= 1.4 * cur) and not (cur < tot / 20000 and c < tot / 20000):
        segs.append((start, idx - 1, cur, acc, accs))
        cur, start, acc, accs = c, idx, 0, 0
    acc += c
    accs += sm
segs.append((start, len(data) - 1, cur, acc, accs))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
for s in segs:
    if s[3] > thr * tot or s[4] > thr * tots:
        print("sass[%4d..%4d] n=%4d per-inst~%10d  inst %5.1f%%  samples %5.1f%%   %s" % (
            s[0], s[1], s[1] - s[0] + 1, s[2], 100 * s[3] / tot, 100 * s[4] / tots, data[s[0]][0][:60]))
if len(sys.argv) > 3:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for i in range(a, b + 1):
        print("%4d %10d %6d  %s" % (i, data[i][1], data[i][2], data[i][0]))
