"""`CausalImpact(data, pre_period, post_period, model=None, model_args={}, alpha=0.05)` -- the
unchanged public entry point (/root/reference/causalimpact/main.py:256-411), with the fit,
forecast and reductions executed by the sm_100a kernels."""
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from . import data as cidata
from . import inferences as inferrer
from . import model as cimodel
from . import plot as plotter
from . import summary as summarizer


class CausalImpact:
    """Bayesian structural time-series estimate of the effect of an intervention.

    Result attributes (same names as the reference, main.py:267-280, 358-359, 380-411): data,
    processed_data_index, pre_period, post_period, pre_data, post_data, alpha, model_args, model,
    normed_pre_data, normed_post_data, observed_time_series, mu_sig, model_samples,
    model_kernel_results, one_step_dist, posterior_dist, inferences, summary_data, p_value."""

    def __init__(self, data, pre_period, post_period, model=None, model_args: Dict[str, Any] = {},
                 alpha: float = 0.05):
        processed = cidata.process_input_data(data, pre_period, post_period, model, model_args, alpha)
        self.data = data
        self.processed_data_index = processed['data'].index
        self.pre_period = processed['pre_period']
        self.post_period = processed['post_period']
        self.pre_data = processed['pre_data']
        self.post_data = processed['post_data']
        self.alpha = processed['alpha']
        self.model_args = processed['model_args']
        self.model = processed['model']
        self.normed_pre_data = processed['normed_pre_data']
        self.normed_post_data = processed['normed_post_data']
        self.observed_time_series = processed['observed_time_series']
        self.mu_sig = processed['mu_sig']
        self._mask = processed['mask']
        self._fit_model()
        self._process_posterior_inferences()
        self._summarize_inferences()

    @staticmethod
    def batch(data, pre_period, post_period, model_args: Dict[str, Any] = {}, alpha: float = 0.05,
              chains: int = 1, seed: int = 0, series_col: Optional[str] = None, time_col: Optional[str] = None,
              inferences: bool = False, max_series_per_pass: Optional[int] = None):
        """Many series in ONE pass on the GPU (the reference can only loop over
        `CausalImpact(...)` calls, each paying its own pandas preprocessing).

        data: array [N, T_total, 1 + k] (column 0 = response, the rest covariates; NaN in the response marks
        missing steps)  |  a sequence of N DataFrames (any mix of indexes; series sharing an index are processed
        as one block)  |  a long-format DataFrame with `series_col` (and `time_col`).  pre_period / post_period:
        `[first, last]` as int positions, date strings or Timestamps, resolved against every index exactly as the
        constructor does (same validation, same error strings; `batch_input.BatchInputError.series` names the
        offending series).  Gaps in a DatetimeIndex become masked steps (data.py:333-334).  model_args / alpha as
        for the constructor.  `inferences=True` also returns the 16-column inference frames
        (`BatchResult.inferences_frame(i)`).  Returns a `BatchResult`."""
        from .batch import causal_impact_batch
        from .batch_input import process_batch_input
        alpha = cidata.process_alpha(alpha)
        args = cimodel.process_model_args(dict(model_args) if model_args else {})
        cimodel.check_capability_limits(args)
        from . import _native as nt
        nt.require_cuda()
        dev = torch.device(f'cuda:{torch.cuda.current_device()}')
        proc = process_batch_input(data, pre_period, post_period, series_col, time_col)
        n = proc['n']
        summ = np.empty((n, 10, 2), dtype=np.float32)
        pval = np.empty((n,), dtype=np.float32)
        infs: List[Any] = [None] * n if inferences else None
        group_of = np.empty((n,), dtype=np.int64)
        for gi, g in enumerate(proc['groups']):
            T, Tf, _ = g.shape_key
            step = max_series_per_pass
            if step is None:          # bound the [n, niter, T] one-step sample tensor of the inference path
                step = max(1, int(6e9 // (4 * args['niter'] * max(T, 1)))) if inferences else len(g.positions)
            for lo in range(0, len(g.positions), step):
                sl = slice(lo, min(lo + step, len(g.positions)))
                y_pre, y_post, X = g.tensors(dev, sl.start, sl.stop)
                res = causal_impact_batch(y_pre, y_post, X,
                                          fit_method=args['fit_method'], nseasons=args['nseasons'],
                                          season_duration=args['season_duration'],
                                          prior_level_sd=args['prior_level_sd'], standardize=args['standardize'],
                                          niter=args['niter'], alpha=alpha, chains=chains, seed=seed + lo,
                                          return_inferences=inferences, device=dev)
                pos = g.positions[sl]
                summ[pos] = res['summary'].cpu().numpy()
                pval[pos] = res['p_value'].cpu().numpy()
                group_of[pos] = gi
                if inferences:
                    arr = res['inferences'].cpu().numpy()
                    present = (~torch.isnan(y_post)).cpu().numpy()
                    for j, p_ in enumerate(pos):
                        infs[p_] = (arr[j], present[j])
        return BatchResult(summ, pval, alpha, args, infs, proc['groups'], group_of, proc['labels'])

    def plot(self, panels: List[str] = ['original', 'pointwise', 'cumulative'], figsize: Tuple[int] = (10, 7),
             show: bool = True) -> None:
        plotter.plot(self.inferences, self.pre_data, self.post_data[self._mask], panels=panels,
                     figsize=figsize, show=show)

    def summary(self, output: str = 'summary', digits: int = 2) -> str:
        if not isinstance(digits, int):
            raise ValueError(f'Input value for digits must be integer. Received "{type(digits)}" instead.')
        return summarizer.summary(self.summary_data, self.p_value, self.alpha, output, digits)

    def _fit_model(self) -> None:
        self.model._Tf = len(self.post_data)
        self.model_samples, self.model_kernel_results = cimodel.fit_model(
            self.model, self.observed_time_series, self.model_args['fit_method'],
            seed=self.model_args.get('seed'))

    def _process_posterior_inferences(self) -> None:
        self.one_step_dist = cimodel.build_one_step_dist(self.model, self.observed_time_series, self.model_samples)
        self.posterior_dist = cimodel.build_posterior_dist(self.model, self.observed_time_series,
                                                           self.model_samples, len(self.post_data))
        self.inferences = inferrer.compile_posterior_inferences(
            self.processed_data_index, self._mask, self.pre_data, self.post_data, self.one_step_dist,
            self.posterior_dist, self.mu_sig, self.alpha, self.model_args['niter'])

    def _summarize_inferences(self) -> None:
        """A fresh set of forecast trajectories -> summary table + tail-area p (main.py:361-386)."""
        niter = self.model_args['niter']
        dev = self.posterior_dist.db.device if hasattr(self.posterior_dist, 'db') else None
        if hasattr(self.posterior_dist, 'sample_device'):
            ms = None if self.mu_sig is None else torch.tensor([[self.mu_sig[0]], [self.mu_sig[1]]],
                                                               dtype=torch.float32, device=dev)
            sims = self.posterior_dist.sample_device(niter, ms)
        else:
            from .misc import maybe_unstandardize
            sims = maybe_unstandardize(np.squeeze(self.posterior_dist.sample(niter).numpy()), self.mu_sig)
        mask = np.asarray(self._mask, dtype=bool)
        sims = sims[:, torch.as_tensor(mask.copy(), device=sims.device)] if torch.is_tensor(sims) else sims[:, mask]
        post_masked = self.post_data[mask]
        self.summary_data = inferrer.summarize_posterior_inferences(
            self.inferences['post_preds_means'], post_masked, sims, self.alpha)
        self.p_value = inferrer.compute_p_value(sims, self.post_data.iloc[:, 0].sum())


class BatchResult:
    """Per-series summary tables, tail-area p-values and (optionally) inference frames of `CausalImpact.batch`."""

    def __init__(self, summary: np.ndarray, p_values: np.ndarray, alpha: float, model_args: Dict[str, Any],
                 inferences=None, groups=None, group_of=None, labels=None):
        self._summary = summary            # [N, 10, 2]
        self.p_values = p_values           # [N]
        self.alpha = alpha
        self.model_args = model_args
        self._inferences = inferences      # per series: ([16, T + Tf + 1] array, post-period mask) or None
        self._groups, self._group_of = groups, group_of
        self.labels = labels               # series labels of a long-format input

    def __len__(self):
        return self._summary.shape[0]

    def summary_data(self, i: int) -> pd.DataFrame:
        """The 10x2 frame of series i (same index / columns as `CausalImpact.summary_data`)."""
        return pd.DataFrame(self._summary[i].astype(np.float64), columns=['average', 'cumulative'],
                            index=inferrer.SUMMARY_INDEX)

    def summary(self, i: int, output: str = 'summary', digits: int = 2) -> str:
        return summarizer.summary(self.summary_data(i), float(self.p_values[i]), self.alpha, output, digits)

    def inferences_frame(self, i: int) -> pd.DataFrame:
        """The 16-column inference frame of series i, indexed and laid out like `CausalImpact.inferences`
        (inferences.py:208-249: masked post steps dropped, cumulative columns prefixed at the last pre label)."""
        if self._inferences is None:
            raise ValueError('CausalImpact.batch was called without inferences=True.')
        out, mask = self._inferences[i]
        g = self._groups[int(self._group_of[i])]
        return inferrer.frame_from_reduction(out.astype(np.float64), mask, g.pre_index, g.post_index, g.range_index)

    def to_frame(self) -> pd.DataFrame:
        """One row per series: the 'average' column of every summary row, the cumulative effect and p."""
        cols = {name: self._summary[:, r, 0] for r, name in enumerate(inferrer.SUMMARY_INDEX)}
        cols['cum_abs_effect'] = self._summary[:, 4, 1]
        cols['p_value'] = self.p_values
        return pd.DataFrame(cols, index=self.labels)
