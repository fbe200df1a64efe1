"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg, order = None, collections.defaultdict(list), []
for r in rows:
    if len(r) > 5 and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        name = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
        try:
            agg[name].append(float(d['Metric Value'].replace(',', '')) * (1e-3 if d['Metric Unit'] in ('ns', 'nsecond') else 1))
        except ValueError:
            pass
tot = sum(sum(v) for v in agg.values())
print(f'{"kernel":50s} {"n":>5s} {"avg us":>9s} {"min":>8s} {"max":>8s} {"share":>7s}')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f'{k[:50]:50s} {len(v):5d} {sum(v)/len(v):9.1f} {min(v):8.1f} {max(v):8.1f} {sum(v)/tot*100:6.1f}%')
print(f'total {tot/1e3:.2f} ms over {sum(len(v) for v in agg.values())} launches')
