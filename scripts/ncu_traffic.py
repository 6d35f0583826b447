"""Regenerate profiles/r2_eval_traffic.json -- the measured DRAM traffic per launch of the evaluation kernel that
bench.py reports as `roofline.traffic` -- from ONE `ncu --set full` capture of scripts/prof_eval.py.  The file is keyed
by a hash of the kernel sources (bench.eval_source_sha): bench.py refuses the value after any kernel change.
usage: python scripts/ncu_traffic.py report.ncu-rep [series chains T k nseasons]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
rep = sys.argv[1]
wl = [int(x) for x in sys.argv[2:7]] if len(sys.argv) >= 7 else [10000, 8, 365, 50, 7]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
best = None
for r in rows[2:]:
    d = dict(zip(h, r))
    if 'k_eval' not in d.get('Kernel Name', ''):
        continue
    u = dict(zip(h, units))
    rd = float(d['dram__bytes_read.sum']) * UNIT[u['dram__bytes_read.sum']]
    wr = float(d['dram__bytes_write.sum']) * UNIT[u['dram__bytes_write.sum']]
    best = {'kernel': d['Kernel Name'].split('(')[0], 'dram_read_bytes': rd, 'dram_write_bytes': wr,
            'traffic_bytes_per_launch': rd + wr, 'gpu_time_us': float(d['gpu__time_duration.sum'])}
assert best, 'no k_eval launch in the report'
best.update({'workload': dict(zip(['series', 'chains', 'T', 'k', 'nseasons'], wl)), 'source_sha': bench.eval_source_sha(),
             'report': os.path.basename(rep), 'how': 'ncu --set full --clock-control none, one launch, dram__bytes_read.sum + dram__bytes_write.sum'})
json.dump(best, open(os.path.join(ROOT, 'profiles', 'r2_eval_traffic.json'), 'w'), indent=1)
print(json.dumps(best))
