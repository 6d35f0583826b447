"""Profile target: the posterior-predictive samplers (K7) at 2000 series x 100 draws x 1000 samples."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from tfcausalimpact_b200.engine import DeviceBatch
N, T, Tf, k, S, D, niter = 2000, 365, 30, 50, 7, 100, 1000
y, X = bench.synth(N, T, Tf, k, S, 5, 'cuda')
db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N * D, device='cuda') * 0.3
u[0] -= 1; u[1] -= 2
mu_sig = torch.stack([torch.zeros(N, device='cuda'), torch.ones(N, device='cuda')])
pm, pv, fs = db.filter_predict(u)
for _ in range(2):
    db.forecast_sample(u, fs, niter, 3, mu_sig)
    db.onestep_sample(pm, pv, niter, 4, mu_sig)
torch.cuda.synchronize()
print('ok')
