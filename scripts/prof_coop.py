"""Profiling driver: a few evaluations of the lane-cooperative kernel at a sub-wave batch.
usage: python scripts/prof_coop.py [series] [tpl] [k]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tfcausalimpact_b200 import _native as nt
from tfcausalimpact_b200.engine import DeviceBatch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
tpl = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k = int(sys.argv[3]) if len(sys.argv) > 3 else 50
G, T, Tf, S = 8, 365, 30, 7
lib = nt.lib()
nt.check(lib.ci_set_option(b'coop_threads_per_lane', ctypes.c_long(tpl)), 'opt')
nt.check(lib.ci_set_option(b'coop_max_lanes', ctypes.c_long(1 << 30)), 'opt')
rng = np.random.default_rng(0)
y = rng.normal(size=(N, T)).astype(np.float32)
X = rng.normal(size=(N, T + Tf, k)).astype(np.float32)
db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N * G, device='cuda') * 0.3
lp = torch.empty(N * G, device='cuda'); gr = torch.empty_like(u)
for _ in range(4):
    db.logp_grad(u, lp, gr)
torch.cuda.synchronize()
print('ok', float(lp.sum()))
