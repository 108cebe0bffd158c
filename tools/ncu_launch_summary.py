"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    k = row['Kernel Name'].replace('<unnamed>::', '').replace('void ', '')[:80]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(a[1] for a in agg.values())
print(f'total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f'{k:80s} {a[0]:6d} {a[1] / 1e3:9.3f} ms {a[1] / a[0]:8.1f} us {a[1] / tot:6.1%}')
