"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for row in csv.DictReader(lines):
    if row['Metric Name'] != 'gpu__time_duration.sum': continue
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1000. if u == 'ns' else v * 1000. if u == 'ms' else v
    name = re.sub(r'\(.*', '', re.sub(r'<.*', '', row['Kernel Name'])).replace('void ', '')
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d} avg={t/n:8.1f}  {k[:80]}")
