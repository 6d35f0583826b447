"""Developer tool: time ci_kalman_logp_grad for every scripts/_variants/*.so on the same box.
usage: python scripts/ab_eval.py [iters] [rounds]"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
from tfcausalimpact_b200 import _native as nt
nt._LIB_PATH = sys.argv[1]
import numpy as np, torch
from tfcausalimpact_b200.engine import DeviceBatch
it = int(sys.argv[2])
N, G, T, Tf, k, S = int(os.environ.get('AB_N', 10000)), 8, 365, 30, int(os.environ.get('AB_K', 50)), 7
rng = np.random.default_rng(0)
y = rng.normal(size=(N, T)).astype(np.float32)
X = rng.normal(size=(N, T + Tf, k)).astype(np.float32)
db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N * G, device='cuda') * 0.3
lp = torch.empty(N * G, device='cuda'); gr = torch.empty_like(u)
for _ in range(5): db.logp_grad(u, lp, gr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(it): db.logp_grad(u, lp, gr)
e1.record(); torch.cuda.synchronize()
ms_eval = e0.elapsed_time(e1) / it
# one HMC transition = 15 evaluations with the fused leapfrog epilogue + the boundary kernel
step0 = torch.full_like(u, 0.01)
uh = u.clone()
draws = torch.empty((10, db.P, N * G), device='cuda')
db.hmc_run(uh, step0, 3, 0, 0, 15, 0.75, seed=1, draws=draws)
torch.cuda.synchronize()
e0.record()
db.hmc_run(uh, step0, 10, 0, 0, 15, 0.75, seed=2, draws=draws)
e1.record(); torch.cuda.synchronize()
print('%%s eval %%.4f ms  hmc transition %%.3f ms  lp_sum=%%.6e grad_abs=%%.6e' %% (os.path.basename(sys.argv[1]), ms_eval,
      e0.elapsed_time(e1) / 10, float(lp.double().sum()), float(gr.double().abs().sum())))
''' % ROOT
it = sys.argv[1] if len(sys.argv) > 1 else '30'
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
libs = sorted(glob.glob(os.path.join(ROOT, 'scripts', '_variants', '*.so')))
for r in range(rounds):
    for lib in libs:
        subprocess.run([sys.executable, '-c', CHILD, lib, it], timeout=120)
