"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
lines = [l for l in open(src) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', re.sub(r'<.*', '', row['Kernel Name'])).replace('void ', '')
    t = float(row['Metric Value'].replace(',', '')) / 1000.0
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
with open(dst, 'w') as f:
    f.write('# %s\n# per-launch times are cold-cache and serialised by ncu: compare shares, not absolutes\n' % title)
    f.write('# total %.1f us over %d launches\n' % (tot, sum(v[0] for v in agg.values())))
    f.write('%8s %12s %7s  %s\n' % ('launches', 'total_us', 'share', 'kernel'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%8d %12.1f %6.1f%%  %s\n' % (v[0], v[1], 100 * v[1] / tot, k))
print(open(dst).read()[:1500])
