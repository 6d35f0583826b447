"""Profile target: a few configs[4]-shaped evaluations (100 lanes x 10,000 steps, nseasons = 52)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from tfcausalimpact_b200.engine import DeviceBatch
N, T, Tf, S = 100, 10000, 520, 52
y, _ = bench.synth(N, T, Tf, 0, S, 5, 'cuda')
db = DeviceBatch(y, None, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N, device='cuda') * 0.3
for _ in range(3):
    lp, g = db.logp_grad(u)
torch.cuda.synchronize()
print('ok', float(lp.sum()))
