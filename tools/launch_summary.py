"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel over the last `steps` steps.
usage: launch_summary.py launches.csv launches_per_step steps"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == 'gpu__time_duration.sum']
per, steps = int(sys.argv[2]), int(sys.argv[3])
rows = rows[-per * steps:]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r'\(.*', '', r[4]).replace('ryk::', '')
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[14].replace(',', '')) / 1e3
tot = sum(a[1] for a in agg.values())
print(f'{"kernel":60s} {"launches/step":>13s} {"us/step":>9s} {"share":>6s}')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:60]:60s} {n / steps:13.1f} {us / steps:9.1f} {us / tot:6.1%}')
print(f'{"TOTAL":60s} {len(rows) / steps:13.1f} {tot / steps:9.1f}')
