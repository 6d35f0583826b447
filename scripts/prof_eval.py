"""Profile target: a few cfg4-shaped evaluations (ci_kalman_logp_grad) and nothing else."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tfcausalimpact_b200.engine import DeviceBatch
N, G, T, Tf, k, S = 10000, 8, 365, 30, int(sys.argv[2]) if len(sys.argv) > 2 else 50, 7
g = torch.Generator(device='cuda').manual_seed(0)
y = torch.randn(N, T, generator=g, device='cuda')
X = torch.randn(N, T + Tf, k, generator=g, device='cuda') if k else None
db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N * G, generator=g, device='cuda') * 0.3
lp = torch.empty(N * G, device='cuda'); gr = torch.empty_like(u)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    db.logp_grad(u, lp, gr)
torch.cuda.synchronize()
print('ok', bool(torch.isfinite(lp).all()))
