"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
    tot[name] += ns; cnt[name] += 1
all_ns = sum(tot.values())
print(f'total {all_ns/1e6:.3f} ms over {sum(cnt.values())} launches')
for name, ns in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
    print(f'{ns/1e6:10.3f} ms {100*ns/all_ns:6.2f}% {cnt[name]:6d}x  {name[:90]}')
