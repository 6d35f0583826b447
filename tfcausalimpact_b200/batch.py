"""Batched front door: N independent series of one shape fitted, forecast and summarised in one
pass on the GPU (and across GPUs with `distributed`), without materialising per-series Python
objects.  This is what BASELINE.json's 1k / 10k-series configurations exercise; the reference can
only express them as N sequential `CausalImpact(...)` calls."""
from typing import Optional

import numpy as np
import torch

from . import distributed as D
from .engine import DeviceBatch, reduce_inferences, reduce_summary

SUMMARY_ROWS = ['actual', 'predicted', 'predicted_lower', 'predicted_upper', 'abs_effect', 'abs_effect_lower',
                'abs_effect_upper', 'rel_effect', 'rel_effect_lower', 'rel_effect_upper']


def standardize_batch(y_pre, y_post, X):
    """Per-series standardisation by PRE-period statistics of every column (data.py:206-229,
    misc.py:27-54: population std, NaN-aware).  Returns standardised tensors and mu_sig [2, N].
    Covariate rows inserted by the gap regularisation are NaN on input and 0 afterwards, and a covariate that is
    constant over the pre-period (a dummy that switches on with the intervention: std 0 -> NaN / inf) becomes 0 as
    well -- the reference zero-fills the standardised design matrix the same way (model.py:303-308)."""
    T = y_pre.shape[1]
    m = torch.nanmean(y_pre, dim=1, keepdim=True)
    sd = torch.sqrt(torch.nanmean((y_pre - m) ** 2, dim=1, keepdim=True))
    ys_pre, ys_post = (y_pre - m) / sd, (y_post - m) / sd
    Xs = None
    if X is not None:
        xm = torch.nanmean(X[:, :T], dim=1, keepdim=True)
        xs = torch.sqrt(torch.nanmean((X[:, :T] - xm) ** 2, dim=1, keepdim=True))
        Xs = torch.nan_to_num((X - xm) / xs, nan=0.0, posinf=0.0, neginf=0.0)
    return ys_pre, ys_post, Xs, torch.cat([m.t(), sd.t()], dim=0).contiguous()


def causal_impact_batch(y_pre, y_post, X=None, fit_method='vi', nseasons=1, season_duration=1,
                        prior_level_sd=0.01, standardize=True, niter=1000, alpha=0.05, chains=1,
                        num_results=100, num_warmup=50, predictive_draws=None, seed=0, device=None,
                        return_inferences=False):
    """y_pre [N,T], y_post [N,Tf] (NaN = missing), X [N,T+Tf,k] or None -> dict with
    'summary' [N,10,2] (rows SUMMARY_ROWS; columns average, cumulative), 'p_value' [N] and the
    posterior draws in unconstrained space; with `return_inferences` also 'inferences' [N,16,T+Tf+1], the
    16 columns of the reference's inference frame (inferences.py:208-246; masked post steps compacted to the
    front, as `ci_reduce_inferences` lays them out).  Under torchrun the series are sharded over ranks and
    the summaries gathered; inputs are then this rank's shard."""
    dev = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
    y_pre = torch.as_tensor(y_pre, dtype=torch.float32, device=dev)
    y_post = torch.as_tensor(y_post, dtype=torch.float32, device=dev)
    X = None if X is None else torch.as_tensor(X, dtype=torch.float32, device=dev)
    N, T = y_pre.shape
    Tf = y_post.shape[1]
    mu_sig = None
    ys_pre, Xs = y_pre, (None if X is None else torch.nan_to_num(X, nan=0.0))
    if standardize:
        ys_pre, _, Xs, mu_sig = standardize_batch(y_pre, y_post, X)
    db = DeviceBatch(ys_pre, Xs, Tf=Tf, nseasons=nseasons, season_duration=season_duration,
                     prior_level_sd=prior_level_sd, device=dev)
    if fit_method == 'vi':
        mu, rho = db.vi_fit(1, 1, 200, 0.1, 0.01, seed)
        u = db.vi_draw(mu, rho, 100, seed + 1)                                  # [P, N*100]
        stats = None
    elif fit_method == 'hmc':
        mu, rho = db.vi_fit(chains, 5, 150, 0.1, 0.01, seed)
        u0 = db.vi_draw(mu, rho, 1, seed + 1)
        step0 = torch.nn.functional.softplus(rho).contiguous()
        draws, stats = db.hmc_run(u0, step0, num_results, num_warmup, int(0.8 * num_warmup), 15, 0.75, seed + 2)
        if predictive_draws is not None and predictive_draws < draws.shape[0]:
            draws = draws[-predictive_draws:]                                        # thin: last draws of every chain
        # lanes of the predictive path = series x (chains x draws)
        R, P, B = draws.shape
        u = draws.view(R, P, N, chains).permute(1, 2, 3, 0).reshape(P, N * chains * R).contiguous()
    else:
        raise ValueError('fit_method can be either "hmc" or "vi".')
    pm, pv, fs = db.filter_predict(u)
    sims, post_mean = db.forecast_sample(u, fs, niter, seed + 3, mu_sig)
    inferences = None
    if return_inferences:
        # a first set of predictive samples feeds the 16-column frame, a FRESH set of forecast trajectories the
        # summary table and p-value -- as the reference does (inferences.py:108-117, main.py:377)
        pre_sims, pre_mean = db.onestep_sample(pm, pv, niter, seed + 4, mu_sig)
        inferences = reduce_inferences(y_pre, y_post, pre_sims, pre_mean, sims, post_mean, alpha)
        del pre_sims
        sims, post_mean = db.forecast_sample(u, fs, niter, seed + 5, mu_sig)
    out = reduce_summary(y_post, post_mean, sims, alpha)                        # [N, 21]
    return {'summary': out[:, :20].reshape(N, 10, 2), 'p_value': out[:, 20], 'u_draws': u,
            'hmc_stats': stats, 'mu_sig': mu_sig, 'engine': db, 'inferences': inferences}


def causal_impact_sharded(y_pre, y_post, X=None, **kw):
    """Host arrays for ALL series in; every rank fits its `shard_range` block and all ranks get
    the gathered [N,10,2] summary and [N] p-values back."""
    rank, world, local = D.init_from_env()
    n = y_pre.shape[0]
    lo, hi = D.shard_range(n, rank, world)
    dev = torch.device(f'cuda:{local}')
    res = causal_impact_batch(y_pre[lo:hi], y_post[lo:hi], None if X is None else X[lo:hi], device=dev, **kw)
    summ = D.gather_series(res['summary'].reshape(hi - lo, 20), n).reshape(n, 10, 2)
    pval = D.gather_series(res['p_value'].reshape(hi - lo, 1), n).reshape(n)
    return {'summary': summ, 'p_value': pval}
