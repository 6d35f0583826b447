"""Developer tool: evaluation / HMC-transition time of the warp-specialised kernel (ci_eval_ws.cu) against the
single-warp thread-per-lane kernel for mid-size batches, on one box.  usage: python scripts/ab_ws.py [k] [chains]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tfcausalimpact_b200 import _native as nt
from tfcausalimpact_b200.engine import DeviceBatch
k = int(sys.argv[1]) if len(sys.argv) > 1 else 50
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T, Tf, S = 365, 30, 7
lib = nt.lib()
if os.environ.get('AB_WSMAX'): nt.check(lib.ci_set_option(b'eval_ws_max_lanes', ctypes.c_long(int(os.environ['AB_WSMAX']))), 'opt')
if os.environ.get('AB_NOCOOP'): nt.check(lib.ci_set_option(b'coop_threads_per_lane', ctypes.c_long(0)), 'opt')
def opt(name, v): nt.check(lib.ci_set_option(name, ctypes.c_long(v)), 'ci_set_option')
def time_eval(db, u, it=20):
    lp = torch.empty(u.shape[1], device='cuda'); gr = torch.empty_like(u)
    for _ in range(3): db.logp_grad(u, lp, gr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): db.logp_grad(u, lp, gr)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it, float(lp.double().sum())
def time_hmc(db, u, n=5):
    step0 = torch.full_like(u, 0.01); uh = u.clone()
    draws = torch.empty((n, db.P, u.shape[1]), device='cuda')
    db.hmc_run(uh, step0, 2, 0, 0, 15, 0.75, seed=1, draws=draws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    db.hmc_run(uh, step0, n, 0, 0, 15, 0.75, seed=2, draws=draws)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
rng = np.random.default_rng(0)
for N in [int(x) for x in os.environ.get('AB_SERIES', '1250,1875,2500,3750,5000,5300,6000,7000').split(',')]:
    y = rng.normal(size=(N, T)).astype(np.float32)
    X = rng.normal(size=(N, T + Tf, k)).astype(np.float32)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    u = torch.randn(db.P, N * G, device='cuda') * 0.3
    row = {'series': N, 'chains': G, 'lanes': N * G, 'k': k}
    for ws in (0, 1):
        opt(b'eval_ws', ws)
        ms, lps = time_eval(db, u)
        row[f'eval_ms_ws{ws}'] = round(ms, 4); row[f'lp_sum_ws{ws}'] = lps
        if os.environ.get('AB_HMC', '1') == '1': row[f'hmc_ms_ws{ws}'] = round(time_hmc(db, u), 3)
    opt(b'eval_ws', 1)
    print(json.dumps(row), flush=True)
    del db
