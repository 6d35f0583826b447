"""Developer tool: evaluation time of the lane-cooperative kernel (8 / 4 sub-threads per lane) against the
thread-per-lane kernel over a range of batch sizes, on one box.
usage: python scripts/ab_coop.py [k] [chains]"""
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


def opt(name, v):
    nt.check(lib.ci_set_option(name, ctypes.c_long(v)), 'ci_set_option')


def time_eval(db, u, it=20):
    lp = torch.empty(u.shape[1], device='cuda'); gr = torch.empty_like(u)
    for _ in range(3): db.logp_grad(u, lp, gr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): db.logp_grad(u, lp, gr)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def time_hmc(db, u, n=5):
    step0 = torch.full_like(u, 0.01)
    uh = u.clone()
    draws = torch.empty((n, db.P, u.shape[1]), device='cuda')
    db.hmc_run(uh, step0, 2, 0, 0, 15, 0.75, seed=1, draws=draws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    db.hmc_run(uh, step0, n, 0, 0, 15, 0.75, seed=2, draws=draws)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rng = np.random.default_rng(0)
for N in [int(x) for x in os.environ.get('AB_SERIES', '16,125,625,1250,2500,5000,10000').split(',')]:
    y = rng.normal(size=(N, T)).astype(np.float32)
    X = rng.normal(size=(N, T + Tf, k)).astype(np.float32)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    u = torch.randn(db.P, N * G, device='cuda') * 0.3
    row = {'series': N, 'chains': G, 'lanes': N * G, 'k': k}
    for tpl, mma in ((0, 1), (8, 1), (4, 0), (4, 1)):
        if tpl == 8 and N * G > 40000:
            continue
        opt(b'coop_threads_per_lane', tpl)
        opt(b'coop_max_lanes', 1 << 30)
        opt(b'coop_mma', mma)
        tag = f'tpl{tpl}' + ('_ffma' if tpl == 4 and not mma else '')
        row[f'eval_ms_{tag}'] = round(time_eval(db, u), 4)
        if os.environ.get('AB_HMC', '1') == '1':
            row[f'hmc_ms_{tag}'] = round(time_hmc(db, u), 3)
    print(json.dumps(row), flush=True)
    del db
