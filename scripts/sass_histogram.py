"""Developer tool: opcode histogram of the shipped library -> profiles/r2_sass_opcodes.txt
(cuobjdump -sass tfcausalimpact_b200/libci_b200.so; run after `python -m tfcausalimpact_b200.build`)."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'tfcausalimpact_b200', 'libci_b200.so')],
                     capture_output=True, text=True).stdout
per, cur, tot = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)', line)
    if m and cur:
        per[cur][m.group(2).split('.')[0]] += 1
        tot[m.group(2)] += 1
key = ['UBLKCP', 'SYNCS', 'LDGSTS', 'FFMA2', 'FADD2', 'FMUL2', 'CCTL', 'MUFU', 'HMMA', 'BAR', 'SHFL']
with open(os.path.join(ROOT, 'profiles', 'r2_sass_opcodes.txt'), 'w') as f:
    f.write('# cuobjdump -sass tfcausalimpact_b200/libci_b200.so (sm_100a), opcode histogram (scripts/sass_histogram.py)\n')
    f.write('# markers: bulk TMA (UBLKCP) + mbarrier transactions (SYNCS.*), cp.async (LDGSTS), packed FP32x2 (FFMA2 / FADD2 /\n'
            '# FMUL2), L2 prefetch (CCTL), special-function unit (MUFU), TF32 tensor-core MMA (HMMA.1688.F32.TF32: the regression\n'
            '# passes of k_eval_coop<S,4>); no UTMALDG / UTC*MMA: no tensor maps, no tcgen05 (see DESIGN.md section 7)\n\n')
    f.write('# marker opcodes with modifiers, whole library\n')
    for op, n in sorted(tot.items()):
        if op.startswith(('SYNCS', 'UBLKCP', 'LDGSTS', 'CCTL', 'UTMA', 'HMMA', 'UTC', 'FFMA2', 'FADD2', 'FMUL2', 'MUFU', 'REDUX')):
            f.write(f'{n:8d} {op}\n')
    f.write('\n# per kernel: ' + ' '.join(key) + ' | total instructions\n')
    for k, c in per.items():
        if sum(c.values()) < 50:
            continue
        f.write(f'{k[:72]:72s} ' + ' '.join(f'{c.get(x, 0):5d}' for x in key) + f' | {sum(c.values())}\n')
    f.write('\n# whole library, full opcodes with modifiers (top 60)\n')
    for op, n in tot.most_common(60):
        f.write(f'{n:8d} {op}\n')
print('written')
