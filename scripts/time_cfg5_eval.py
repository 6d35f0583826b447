"""Developer timing of the large-latent evaluation at configs[4] shape (100 lanes x 10,000 steps, nseasons 52).
usage: python scripts/time_cfg5_eval.py [lib.so]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tfcausalimpact_b200 import _native as nt
if len(sys.argv) > 1:
    nt._LIB_PATH = sys.argv[1]
import torch, bench
from tfcausalimpact_b200.engine import DeviceBatch
N, T, Tf, S = 100, 10000, 520, 52
y, _ = bench.synth(N, T, Tf, 0, S, 5, 'cuda')
db = DeviceBatch(y, None, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N, device='cuda') * 0.3
lp, g = db.logp_grad(u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    db.logp_grad(u, lp, g)
e1.record(); torch.cuda.synchronize()
print(os.path.basename(nt._LIB_PATH), 'cfg5 eval %.3f ms' % (e0.elapsed_time(e1) / 5), float(lp.sum()), float(g.abs().sum()))
