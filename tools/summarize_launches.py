"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
idx = {h: i for i, h in enumerate(rows[start])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    try:
        name, v, unit = r[idx['Kernel Name']], float(r[idx['Metric Value']]), r[idx['Metric Unit']]
    except Exception:
        continue
    v = v / 1e6 if unit.startswith('n') else (v / 1e3 if unit.startswith('u') else v)
    name = re.sub(r'\(.*', '', name).replace('void ', '')
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f'{"kernel":58s} {"launches":>8s} {"ms":>10s} {"share":>7s}')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:58s} {v[0]:8d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}%')
print(f'{"total":58s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}')
