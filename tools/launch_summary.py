"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals + one decode step."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
order = []
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    t = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    t *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1)
    order.append((name, t, row.get('Grid Size', ''), row.get('Block Size', '')))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.1f} us over {len(order)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%  {v[0]:5d}x  avg {v[1]/v[0]/1e3:8.2f} us  {k[:80]}")
idx = [i for i, o in enumerate(order) if 'dec_embed' in o[0]]
if len(idx) >= 2:
    step = order[idx[-2]:idx[-1]]
    print(f"one decode step: {sum(o[1] for o in step)/1e3:.1f} us, {len(step)} launches")
    sa = collections.OrderedDict()
    for o in step:
        a = sa.setdefault(o[0], [0, 0.0]); a[0] += 1; a[1] += o[1]
    for k, v in sorted(sa.items(), key=lambda kv: -kv[1][1]):
        print(f"   {v[1]/1e3:8.1f} us {v[0]:3d}x avg {v[1]/v[0]/1e3:7.2f}  {k[:70]}")
    if '-v' in sys.argv:
        for o in step[:16]:
            print(f"      {o[1]/1e3:8.2f} us {o[2]} {o[3]} {o[0][:60]}")
