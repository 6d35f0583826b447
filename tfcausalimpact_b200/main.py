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
              chains: int = 1, seed: int = 0):
        """Many same-shaped series in ONE pass on the GPU (the reference can only
        loop over `CausalImpact(...)` calls).

        data: array [N, T_total, 1 + k] (column 0 = response, the rest covariates; NaN in the response
        marks missing steps); pre_period / post_period: integer [first, last] positions, inclusive,
        shared by all series; model_args / alpha as for the constructor (same validation and errors).
        Returns a `BatchResult`."""
        from .batch import causal_impact_batch
        arr = np.asarray(data, dtype=np.float32)
        if arr.ndim != 3 or arr.shape[0] < 1:
            raise ValueError('data must be an array of shape [N, T, 1 + k].')
        alpha = cidata.process_alpha(alpha)
        args = cimodel.process_model_args(dict(model_args) if model_args else {})
        frame = pd.DataFrame(arr[0])
        pre, post = cidata.process_period(pre_period, frame), cidata.process_period(post_period, frame)
        if pre[1] > post[0]:
            raise ValueError('Values in training data cannot be present in the post-intervention '
                             'data. Please fix your pre_period value to cover at most one point less '
                             'from when the intervention happened.')
        if pre[1] - pre[0] < 3:
            raise ValueError('pre_period must span at least 3 time points.')
        if post[1] < post[0]:
            raise ValueError('post_period last number must be bigger than its first.')
        if post[0] <= pre[1]:
            raise ValueError(f'post_period first value ({post_period[0]}) must '
                             'be bigger than the second value of pre_period '
                             f'({pre_period[1]}).')
        if arr.shape[2] > 1 and np.isnan(arr[:, :, 1:]).any():
            raise ValueError('Input data cannot have NAN values.')
        y_pre, y_post = arr[:, pre[0]:pre[1] + 1, 0], arr[:, post[0]:post[1] + 1, 0]
        X = None
        if arr.shape[2] > 1:
            X = np.concatenate([arr[:, pre[0]:pre[1] + 1, 1:], arr[:, post[0]:post[1] + 1, 1:]], axis=1)
        res = causal_impact_batch(np.ascontiguousarray(y_pre), np.ascontiguousarray(y_post), X,
                                  fit_method=args['fit_method'], nseasons=args['nseasons'],
                                  season_duration=args['season_duration'], prior_level_sd=args['prior_level_sd'],
                                  standardize=args['standardize'], niter=args['niter'], alpha=alpha, chains=chains,
                                  seed=seed)
        return BatchResult(res['summary'].cpu().numpy(), res['p_value'].cpu().numpy(), alpha, args)

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
    """Per-series summary tables and tail-area p-values of `CausalImpact.batch`."""

    def __init__(self, summary: np.ndarray, p_values: np.ndarray, alpha: float, model_args: Dict[str, Any]):
        self._summary = summary            # [N, 10, 2]
        self.p_values = p_values           # [N]
        self.alpha = alpha
        self.model_args = model_args

    def __len__(self):
        return self._summary.shape[0]

    def summary_data(self, i: int) -> pd.DataFrame:
        """The 10x2 frame of series i (same index / columns as `CausalImpact.summary_data`)."""
        return pd.DataFrame(self._summary[i].astype(np.float64), columns=['average', 'cumulative'],
                            index=inferrer.SUMMARY_INDEX)

    def summary(self, i: int, output: str = 'summary', digits: int = 2) -> str:
        return summarizer.summary(self.summary_data(i), float(self.p_values[i]), self.alpha, output, digits)

    def to_frame(self) -> pd.DataFrame:
        """One row per series: the 'average' column of every summary row, the cumulative effect and p."""
        cols = {name: self._summary[:, r, 0] for r, name in enumerate(inferrer.SUMMARY_INDEX)}
        cols['cum_abs_effect'] = self._summary[:, 4, 1]
        cols['p_value'] = self.p_values
        return pd.DataFrame(cols)
