"""Timing of the predictive + reduction kernels (K6-K8) at configs[3] scale."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from tfcausalimpact_b200.engine import DeviceBatch, reduce_summary, reduce_inferences
N, T, Tf, k, S, D, niter = int(os.environ.get('NSER', 10000)), int(os.environ.get('TPRE', 365)), 30, 50, 7, 100, 1000
y, X = bench.synth(N, T, Tf, k, S, 5, 'cuda')
db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N * D, device='cuda') * 0.3
u[0] -= 1; u[1] -= 2
ypost = torch.randn(N, Tf, device='cuda')
mu_sig = torch.stack([torch.zeros(N, device='cuda'), torch.ones(N, device='cuda')])
def timed(name, fn, bytes_moved=None, reps=5):
    # outputs are torch allocations: warm the caching allocator first (a fresh 3 GB cudaMalloc inside the timed region
    # costs milliseconds) and never hold two generations of outputs at once
    out = None
    for _ in range(3):
        out = None
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = None
        out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rec = {'kernel': name, 'ms': ms}
    if bytes_moved: rec['GBps'] = bytes_moved / ms / 1e6; rec['frac_hbm_6537'] = rec['GBps'] / 6536.7
    print(json.dumps(rec), flush=True)
    return out
B = N * D
import ctypes
from tfcausalimpact_b200 import _native as nt
def opt(name, v): nt.check(nt.lib().ci_set_option(name, ctypes.c_long(v)), 'opt')
if os.environ.get('AB'):
    for v, tag in ((0, 'three kernels'), (1, 'fused')):
        opt(b'filter_fused', v)
        timed(f'ci_filter_predict [{tag}]', lambda: db.filter_predict(u), bytes_moved=2 * T * B * 4 + db.X.numel() * 4)
pm, pv, fs = timed('ci_filter_predict (1M lanes x 365)', lambda: db.filter_predict(u), bytes_moved=2 * T * B * 4 + db.X.numel() * 4)
sims_post, mean_post = timed('ci_forecast_sample (N x 1000 x 30)', lambda: db.forecast_sample(u, fs, niter, 3, mu_sig), bytes_moved=N * niter * Tf * 4)
npre = min(N, 2000)   # pre-period sample tensor of 2000 series = 2.9 GB
db2 = DeviceBatch(y[:npre], X[:npre], Tf=Tf, nseasons=S)
pm2, pv2 = pm[:, :npre * D].contiguous(), pv[:, :npre * D].contiguous()
if os.environ.get('AB'):
    opt(b'onestep_tiled', 0)
    timed(f'ci_onestep_sample [sample-per-thread gather] ({npre} x 1000 x 365)', lambda: db2.onestep_sample(pm2, pv2, niter, 4, mu_sig[:, :npre].contiguous()), bytes_moved=npre * niter * T * 4)
    opt(b'onestep_tiled', 1)
sims_pre, mean_pre = timed(f'ci_onestep_sample ({npre} x 1000 x 365)', lambda: db2.onestep_sample(pm2, pv2, niter, 4, mu_sig[:, :npre].contiguous()), bytes_moved=npre * niter * T * 4)
timed('ci_reduce_summary (N x 1000 x 30)', lambda: reduce_summary(ypost, mean_post, sims_post, 0.05), bytes_moved=N * niter * Tf * 4)
timed(f'ci_reduce_inferences ({npre} series)', lambda: reduce_inferences(y[:npre].contiguous(), ypost[:npre].contiguous(), sims_pre, mean_pre, sims_post[:npre].contiguous(), mean_post[:npre].contiguous(), 0.05),
      bytes_moved=npre * niter * (T + 3 * Tf) * 4 + 2 * npre * niter * Tf * 4)
