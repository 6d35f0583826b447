import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from tfcausalimpact_b200.engine import DeviceBatch, reduce_summary
N, T, Tf, S = 100, 10000, 520, 52
y, _ = bench.synth(N, T, Tf, 0, S, 5, 'cuda')
def tick(msg, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print(f'{msg}: {t1 - t0:.3f} s', flush=True); return t1
t0 = time.perf_counter()
db = DeviceBatch(y, None, Tf=Tf, nseasons=S); t0 = tick('setup', t0)
u = torch.randn(db.P, N, device='cuda') * 0.3
lp, g = db.logp_grad(u); t0 = tick('1 eval (100 lanes)', t0)
lp, g = db.logp_grad(u); t0 = tick('1 eval (100 lanes) again', t0)
mu, rho = db.vi_fit(1, 1, 20, 0.1, 0.01, 1); t0 = tick('20 VI steps', t0)
ud = db.vi_draw(mu, rho, 100, 2); t0 = tick('vi_draw', t0)
pm, pv, fs = db.filter_predict(ud); t0 = tick('filter 10k lanes', t0)
sims, mean = db.forecast_sample(ud, fs, 1000, 3); t0 = tick('forecast sample', t0)
ypost = torch.randn(N, Tf, device='cuda')
out = reduce_summary(ypost, mean, sims, 0.05); t0 = tick('reduce_summary', t0)
