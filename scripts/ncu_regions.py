"""Developer tool: split a kernel's executed instructions / stall samples into contiguous SASS regions
(split wherever the per-instruction execution count changes by more than 2x) from `ncu --page source --csv`.
usage: ncu -i rep --page source --csv --print-source sass > src.csv; python scripts/ncu_regions.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, isrc, ismp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
data = [(int(r[ia], 16), r[isrc].strip(), int(r[ismp]), int(r[iex])) for r in rows[2:] if len(r) > iex]
base = data[0][0]
tot_i = sum(d[3] for d in data); tot_s = sum(d[2] for d in data)
regs = []; cur = None
for a, s, smp, ex in data:
    if cur is None or not (cur['ex'] / 1.5 <= max(ex, 1) <= cur['ex'] * 1.5):
        cur = {'lo': a - base, 'ex': max(ex, 1), 'n': 0, 'inst': 0, 'smp': 0}
        regs.append(cur)
    cur['n'] += 1; cur['inst'] += ex; cur['smp'] += smp; cur['hi'] = a - base
print(f'total inst {tot_i} samples {tot_s}')
for r in regs:
    if r['inst'] > 0.004 * tot_i or r['smp'] > 0.004 * tot_s:
        print(f"{r['lo']:6x}-{r['hi']:6x} n={r['n']:5d} exec/instr={r['ex']:9d} inst={100*r['inst']/tot_i:5.1f}% samples={100*r['smp']/tot_s:5.1f}%")
