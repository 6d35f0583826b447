"""Developer tool: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table).
usage: python scripts/launch_summary.py launches.csv "<command that was profiled>" > profiles/rN_launches_summary.md"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6, 'second': 1e3}.get(r[iu], 1e-6)
    a = agg.setdefault(r[ik].split('(')[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'Command: `{sys.argv[2]}` on one B200 (cold-cache, serialised by the profiler: compare SHARES, not absolute times).\n')
print('| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|')
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'| `{k[:90]}` | {n} | {ms:.2f} | {1e3 * ms / n:.1f} | {ms / tot:.3f} |')
