"""Wall-clock split of `CausalImpact.batch` on N DataFrames (daily DatetimeIndex with a gap per period, k covariates):
host preprocessing (`batch_input.process_batch_input`) vs everything else (upload, fit, forecast, summary).
usage: python scripts/time_batch_frontdoor.py [N] [k] [fit_method] [chains]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pandas as pd, torch
from tfcausalimpact_b200 import CausalImpact
from tfcausalimpact_b200 import batch_input

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
method = sys.argv[3] if len(sys.argv) > 3 else 'vi'
chains = int(sys.argv[4]) if len(sys.argv) > 4 else 1
T, Tf, S = 365, 30, 7
rng = np.random.default_rng(0)
full = pd.date_range('2021-01-01', periods=T + Tf, freq='D')
keep = np.ones(T + Tf, dtype=bool)
keep[[100, T + 7]] = False
idx = pd.DatetimeIndex(full[keep].values)
vals = rng.normal(size=(N, T + Tf, 1 + k)).astype(np.float32)
vals[:, :, 0] += 0.3 * vals[:, :, 1:6].sum(axis=2) + np.sin(np.arange(T + Tf) * 2 * np.pi / 7)[None, :]
vals[:, T:, 0] += 0.5
frames = [pd.DataFrame(vals[n][keep], index=idx) for n in range(N)]
pre, post = [full[0], full[T - 1]], [full[T], full[-1]]
args = {'nseasons': S, 'fit_method': method, 'niter': 1000}
CausalImpact.batch(frames[:64], pre, post, model_args=args, chains=chains)          # warm-up (module load, allocator)
torch.cuda.synchronize()
host = {'t': 0.0}
orig = batch_input.process_batch_input


def timed(*a, **kw):
    t0 = time.perf_counter()
    out = orig(*a, **kw)
    host['t'] += time.perf_counter() - t0
    return out


batch_input.process_batch_input = timed
orig_tensors = batch_input.SeriesGroup.tensors


def timed_tensors(self, *a, **kw):      # host side of the upload + device-side validation (includes its tiny syncs)
    t0 = time.perf_counter()
    out = orig_tensors(self, *a, **kw)
    host['t'] += time.perf_counter() - t0
    return out


batch_input.SeriesGroup.tensors = timed_tensors
t0 = time.perf_counter()
res = CausalImpact.batch(frames, pre, post, model_args=args, chains=chains)
torch.cuda.synchronize()
total = time.perf_counter() - t0
print(json.dumps({'series': N, 'k': k, 'T': T, 'Tf': Tf, 'fit_method': method, 'chains': chains, 'total_s': round(total, 3),
                  'host_preprocess_s': round(host['t'], 3), 'host_frac': round(host['t'] / total, 3),
                  'series_per_s': round(N / total, 1), 'abs_effect_mean': float(res.to_frame()['abs_effect'].mean()),
                  'masked_steps_per_series': 2}))
