"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of each kernel in the step."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[ix['Kernel Name']].split('(')[0]
    v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(unit, 1.0)
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'mean us':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {v / cnt[k]:10.2f} {100 * v / s:6.1f}%")
