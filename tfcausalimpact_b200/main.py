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
