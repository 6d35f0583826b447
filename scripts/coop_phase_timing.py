"""Developer tool: per-warp phase durations (globaltimer, ns) of the lane-cooperative kernel from a -DCI_COOP_TIMING=<print every n-th block>
build (python scripts/build_variant.py timing -DCI_COOP_TIMING=1).  usage: python scripts/coop_phase_timing.py series tpl mma [k]"""
import ctypes, os, sys
ROOT = '/root/repo'
sys.path.insert(0, ROOT)
from tfcausalimpact_b200 import _native as nt
nt._LIB_PATH = os.path.join(ROOT, 'scripts', '_variants', 'timing.so')
import numpy as np, torch
from tfcausalimpact_b200.engine import DeviceBatch
N, tpl, mma = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
G, T, Tf, S, k = 8, 365, 30, 7, int(sys.argv[4]) if len(sys.argv) > 4 else 50
lib = nt.lib()
for nme, v in ((b'coop_threads_per_lane', tpl), (b'coop_max_lanes', 1 << 30), (b'coop_mma', mma)):
    nt.check(lib.ci_set_option(nme, ctypes.c_long(v)), 'opt')
rng = np.random.default_rng(0)
y = rng.normal(size=(N, T)).astype(np.float32)
X = rng.normal(size=(N, T + Tf, k)).astype(np.float32)
db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N * G, device='cuda') * 0.3
lp = torch.empty(N * G, device='cuda'); gr = torch.empty_like(u)
db.logp_grad(u, lp, gr); torch.cuda.synchronize()
print('--- second launch', flush=True)
db.logp_grad(u, lp, gr); torch.cuda.synchronize()
