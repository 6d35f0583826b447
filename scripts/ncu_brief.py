"""Developer tool: print the handful of ncu raw-page metrics the DESIGN/profiles summaries quote.
usage: python scripts/ncu_brief.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active', 'sm__inst_executed_pipe_fma',
        'sm__pipe_fma_cycles_active', 'sm__pipe_fmaheavy', 'sm__inst_executed_pipe_lsu', 'l1tex__data_pipe_lsu_wavefronts.avg.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate', 'smsp__average_warps_issue_stalled',
        'smsp__average_warp_latency_issue_stalled', 'sm__cycles_elapsed.max', 'launch__waves', 'smsp__inst_executed_pipe',
        'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_xu', 'smsp__cycles_active.avg']
for r in rows[2:]:
    d = dict(zip(h, r))
    print('==', d.get('Kernel Name', '')[:80])
    for k in h:
        if any(k.startswith(p) for p in KEYS) and 'peak_sustained' not in k.replace('pct_of_peak_sustained_active', '') and '.max' not in k.replace('elapsed.max', '') and '.min' not in k:
            v = d[k]
            try:
                if float(v) == 0: continue
            except ValueError:
                pass
            print(f'  {k} = {v}')
