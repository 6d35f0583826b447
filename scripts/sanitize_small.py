"""Small pass over every kernel family for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from helpers import synth_problem
from tfcausalimpact_b200.batch import causal_impact_batch
for (N, T, Tf, k, S, method, chains) in [(5, 41, 7, 6, 7, 'hmc', 3), (3, 30, 5, 0, 1, 'vi', 1), (2, 37, 6, 2, 9, 'vi', 1),
                                          (3, 45, 9, 5, 4, 'hmc', 2)]:
    y_pre, y_post, X = synth_problem(N, T, Tf, k, S, 1, seed=3, standardize=False, nan_frac=0.05)
    res = causal_impact_batch(y_pre, y_post, X if k else None, fit_method=method, nseasons=S, niter=300, chains=chains,
                              num_results=5, num_warmup=5, seed=2)
    torch.cuda.synchronize()
    print(N, T, k, S, method, bool(torch.isfinite(res['summary']).all()))
print('done')
