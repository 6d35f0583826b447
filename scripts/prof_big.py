import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from tfcausalimpact_b200.engine import DeviceBatch
N, T, Tf, S = 100, 1000, 52, 52
y, _ = bench.synth(N, T, Tf, 0, S, 5, 'cuda')
db = DeviceBatch(y, None, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N, device='cuda') * 0.3
for _ in range(3):
    lp, g = db.logp_grad(u)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    lp, g = db.logp_grad(u)
torch.cuda.synchronize()
print('ok', bool(torch.isfinite(lp).all()), 'ms/eval', (time.perf_counter() - t0) / 5 * 1e3)
