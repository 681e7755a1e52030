"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mu = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split('(')[0][:80]
    v = float(r[mv].replace(',', ''))
    v = v / 1e3 if r[mu] == 'ns' else (v * 1e3 if r[mu] == 'ms' else v)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.0f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]:10.0f} us {100*v[1]/tot:5.1f}% n={v[0]:4d} avg {v[1]/v[0]:8.1f} us  {k}")
