"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[i]
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[i + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0][:72]
    v = float(r[mv].replace(',', ''))
    v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}.get(r[mu], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print('| kernel | launches | total ms | share |\n|---|---|---|---|')
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print('| %s | %d | %.3f | %.1f%% |' % (n, a[0], a[1], 100 * a[1] / tot))
print('| total | | %.3f | |' % tot)
