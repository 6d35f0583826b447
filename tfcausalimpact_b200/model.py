"""Model recipe, fit and predictive distributions of the front door.

Mirrors the public surface of /root/reference/causalimpact/model.py (process_model_args :42,
check_input_model :137, build_default_model :239, fit_model :324, build_one_step_dist :393,
build_posterior_dist :421) with tensorflow-probability removed: the structural time-series recipe
is a light spec object lowered to the fixed state-space layout of the sm_100a kernels
(`engine.DeviceBatch`), and every numeric step runs on the GPU through the C-ABI.  There is no CPU
fallback: without the CUDA library these functions raise.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd
import torch

kLocalLevelPriorSampleSize = 32   # model.py:39
_VI_STEPS = 200                   # model.py:369
_VI_DRAWS = 100                   # model.py:384
# tfp.sts.fit_with_hmc defaults (the reference passes none, model.py:361-364)
_HMC = dict(num_results=100, num_warmup_steps=50, num_leapfrog_steps=15, num_variational_steps=150,
            variational_sample_size=5)


def process_model_args(model_args: Dict[str, Any]) -> Dict[str, Any]:
    """Fill defaults and type-check the six public keys (same messages as model.py:102-133).
    Extra optional keys understood by this implementation: `seed` (int)."""
    standardize = model_args.get('standardize', True)
    if not isinstance(standardize, bool):
        raise ValueError('standardize argument must be of type bool.')
    model_args['standardize'] = standardize
    prior_level_sd = model_args.get('prior_level_sd', 0.01)
    if not isinstance(prior_level_sd, float):
        raise ValueError('prior_level_sd argument must be of type float.')
    model_args['prior_level_sd'] = prior_level_sd
    niter = model_args.get('niter', 1000)
    if not isinstance(niter, int):
        raise ValueError('niter argument must be of type int.')
    model_args['niter'] = niter
    fit_method = model_args.get('fit_method', 'vi')
    if fit_method not in {'hmc', 'vi'}:
        raise ValueError('fit_method can be either "hmc" or "vi".')
    model_args['fit_method'] = fit_method
    nseasons = model_args.get('nseasons', 1)
    if not isinstance(nseasons, int):
        raise ValueError('nseasons argument must be of type int.')
    model_args['nseasons'] = nseasons
    season_duration = model_args.get('season_duration', 1)
    if not isinstance(season_duration, int):
        raise ValueError('season_duration argument must be of type int.')
    if nseasons <= 1 and season_duration > 1:
        raise ValueError('nseasons must be bigger than 1 when season_duration is also '
                         'bigger than 1.')
    model_args['season_duration'] = season_duration
    return model_args


MAX_NITER = 8192          # draws one sorting CTA of ci_reduce_inferences / ci_reduce_summary holds (csrc/ci_reduce.cu)
MAX_LATENT = 96           # latent size (1 + nseasons [+ trend slope]) of the shared-memory filter (csrc/ci_bigd.cu)


def check_capability_limits(model_args: Dict[str, Any]) -> None:
    """Limits of this implementation that the reference does not have, reported BEFORE any fitting with a clear
    message (the C-ABI would return CI_ERR_UNSUPPORTED only after the fit has run)."""
    if model_args['niter'] > MAX_NITER:
        raise ValueError(f'niter={model_args["niter"]} exceeds the {MAX_NITER} posterior-predictive samples '
                         'this implementation supports.')
    if model_args['niter'] < 1:
        raise ValueError('niter must be at least 1.')
    if 1 + model_args['nseasons'] > MAX_LATENT:
        raise ValueError(f'nseasons={model_args["nseasons"]} exceeds the {MAX_LATENT - 1} seasons this '
                         'implementation supports (latent size 1 + nseasons <= 96).')


# ------------------------------------------------------------------------------------------------
# Spec objects (attribute names follow what the reference's tests / notebook introspect:
# model.components[i], .design_matrix.to_dense(), .num_seasons, .num_steps_per_season,
# .parameters[i].prior -- tests/test_main.py:59-64,188-191, tests/test_model.py:163-220)
# ------------------------------------------------------------------------------------------------
@dataclass
class Prior:
    """Descriptor of a prior density (the kernels hard-wire the arithmetic)."""
    family: str
    params: Dict[str, float]
    dtype: Any = np.float32

    def __repr__(self):
        args = ', '.join(f'{k}={v:.6g}' for k, v in self.params.items())
        return f'{self.family}({args})'


@dataclass
class Parameter:
    name: str
    prior: Prior
    bijector: str
    shape: tuple = ()


class DesignMatrix:
    def __init__(self, values):
        self._values = np.ascontiguousarray(np.asarray(values, dtype=np.float32))
        self.shape = self._values.shape
        self.dtype = np.float32

    def to_dense(self):
        return self._values


def build_inv_gamma_sd_prior(sigma_guess: float) -> Prior:
    """V ~ InverseGamma(a = 32/2, b = 32 sigma^2 / 2)  (model.py:196-216)."""
    a = np.float32(kLocalLevelPriorSampleSize / 2)
    b = np.float32(kLocalLevelPriorSampleSize * sigma_guess ** 2 / 2)
    return Prior('InverseGamma', {'concentration': float(a), 'scale': float(b)})


def build_bijector(dist: Prior) -> Prior:
    """sigma = sqrt(V)  (model.py:219-236: TransformedDistribution(dist, Power(.5)))."""
    return Prior('Sqrt' + dist.family, dict(dist.params))


class Component:
    name = 'Component'
    parameters: List[Parameter] = []


class LocalLevel(Component):
    def __init__(self, level_scale_prior: Optional[Prior] = None, observed_time_series=None, name='LocalLevel'):
        self.name = name
        if level_scale_prior is None:
            # tfp.sts.LocalLevel default: LogNormal(log(.05 sd), 3)
            level_scale_prior = Prior('LogNormal', {'loc': float('nan'), 'scale': 3.0})
        self.level_scale_prior = level_scale_prior
        self.parameters = [Parameter('level_scale', level_scale_prior, 'Scale(sd_y) o Softplus')]


class LocalLinearTrend(Component):
    """Level + slope random walks (tfp.sts.LocalLinearTrend with its default priors: both scales
    LogNormal(log(.05 sd), 3), initial slope N(0, sd^2)) -- the trend of the custom-model examples in
    the reference's docs (causalimpact/main.py:191,217,246)."""

    def __init__(self, observed_time_series=None, name='LocalLinearTrend'):
        self.name = name
        ln = Prior('LogNormal', {'loc': float('nan'), 'scale': 3.0})
        self.parameters = [Parameter('level_scale', ln, 'Scale(sd_y) o Softplus'),
                           Parameter('slope_scale', ln, 'Scale(sd_y) o Softplus')]


class LinearRegression(Component):
    """Dense regression with tfp.sts.LinearRegression's default prior: weights ~ StudentT(5, 0, 10)."""

    def __init__(self, design_matrix, name='LinearRegression'):
        self.name = name
        if isinstance(design_matrix, pd.DataFrame):
            design_matrix = design_matrix.values
        self.design_matrix = design_matrix if isinstance(design_matrix, DesignMatrix) else DesignMatrix(design_matrix)
        k = self.design_matrix.shape[1]
        self.parameters = [Parameter('weights', Prior('StudentT', {'df': 5.0, 'loc': 0.0, 'scale': 10.0}), 'Identity', (k,))]


class SparseLinearRegression(Component):
    """Horseshoe-prior regression (tfp.sts.SparseLinearRegression; weights formula pinned by the
    reference's tests/test_main.py:326-339)."""

    def __init__(self, design_matrix, weights_prior_scale=0.1, name='SparseLinearRegression'):
        self.name = name
        if isinstance(design_matrix, pd.DataFrame):
            design_matrix = design_matrix.values
        self.design_matrix = design_matrix if isinstance(design_matrix, DesignMatrix) else DesignMatrix(design_matrix)
        if abs(weights_prior_scale - 0.1) > 1e-12:
            raise ValueError('SparseLinearRegression: only weights_prior_scale=0.1 (the TFP default) is supported.')
        k = self.design_matrix.shape[1]
        ig, hn = Prior('InverseGamma', {'concentration': 0.5, 'scale': 0.5}), Prior('HalfNormal', {'scale': 1.0})
        self.parameters = [
            Parameter('global_scale_variance', ig, 'Softplus'),
            Parameter('global_scale_noncentered', hn, 'Softplus'),
            Parameter('local_scale_variances', ig, 'Softplus', (k,)),
            Parameter('local_scales_noncentered', hn, 'Softplus', (k,)),
            Parameter('weights_noncentered', Prior('Normal', {'loc': 0.0, 'scale': 1.0}), 'Identity', (k,)),
        ]


class Seasonal(Component):
    def __init__(self, num_seasons, num_steps_per_season=1, observed_time_series=None, name='Seasonal'):
        self.name = name
        self.num_seasons = int(num_seasons)
        self.num_steps_per_season = int(num_steps_per_season)
        self.parameters = [Parameter('drift_scale', Prior('LogNormal', {'loc': float('nan'), 'scale': 3.0}),
                                     'Scale(sd_y) o Softplus')]


class Sum:
    """Additive structural time-series model the kernels implement: one trend (LocalLevel |
    LocalLinearTrend) [+ one regression (SparseLinearRegression | LinearRegression)] [+ up to two Seasonal
    components, e.g. the notebook's Month + Year (getting_started.ipynb cell 116)] + observation noise.  Built without `observation_noise_scale_prior` it takes tfp.sts.Sum's default,
    LogNormal(log(.01 sd), 2)."""

    def __init__(self, components, observed_time_series=None, observation_noise_scale_prior: Optional[Prior] = None,
                 prior_level_sd: Optional[float] = None, name='Sum'):
        self.name = name
        self.components = list(components)
        self.observed_time_series = observed_time_series
        self._validate_components()
        if observation_noise_scale_prior is None:
            observation_noise_scale_prior = Prior('LogNormal', {'loc': float('nan'), 'scale': 2.0})
        self.observation_noise_scale_prior = observation_noise_scale_prior
        self.parameters = [Parameter('observation_noise_scale', observation_noise_scale_prior,
                                     'Scale(sd_y) o Softplus')]
        for c in self.components:
            for p in c.parameters:
                self.parameters.append(Parameter(f'{c.name}/_{p.name}', p.prior, p.bijector, p.shape))
        self._engine = None

    def _validate_components(self):
        kinds = [type(c) for c in self.components]
        for c in self.components:
            if not isinstance(c, (LocalLevel, LocalLinearTrend, SparseLinearRegression, LinearRegression, Seasonal)):
                raise ValueError(
                    f'Unsupported structural component {type(c).__name__!r}: the B200 kernels implement '
                    'LocalLevel | LocalLinearTrend, SparseLinearRegression | LinearRegression and Seasonal '
                    'only; other tfp.sts components are rejected (no fallback).')
        if kinds.count(LocalLevel) + kinds.count(LocalLinearTrend) != 1 or \
                kinds.count(SparseLinearRegression) + kinds.count(LinearRegression) > 1 or kinds.count(Seasonal) > 2:
            raise ValueError('Supported model = exactly one trend (LocalLevel or LocalLinearTrend), at most one '
                             'regression (SparseLinearRegression or LinearRegression) and at most two Seasonal '
                             'components.')
        seas = [c for c in self.components if isinstance(c, Seasonal)]
        if len(seas) == 2 and (seas[0].num_seasons <= 1 or seas[1].num_seasons <= 1):
            raise ValueError('Each of two Seasonal components needs num_seasons > 1.')
        if len(seas) == 2 and seas[0].name == seas[1].name:
            raise ValueError('Two Seasonal components need distinct names (their parameters are keyed by name).')
        order = {LocalLevel: 0, LocalLinearTrend: 0, SparseLinearRegression: 1, LinearRegression: 1, Seasonal: 2}
        if [order[k] for k in kinds] != sorted(order[k] for k in kinds):
            raise ValueError('Components must be ordered trend, regression, seasonal (the parameter order of the '
                             'kernels).')

    # ---- lowering to the device layout ----------------------------------------------------------
    def _component(self, cls):
        for c in self.components:
            if isinstance(c, cls):
                return c
        return None

    @property
    def prior_level_sd(self) -> float:
        ll = self._component(LocalLevel)
        if ll is None or ll.level_scale_prior.family == 'LogNormal':
            return 0.01                     # unused: the level scale prior is LogNormal(log(.05 sd), 3)
        pr = ll.level_scale_prior
        if pr.family != 'SqrtInverseGamma' or abs(pr.params['concentration'] - kLocalLevelPriorSampleSize / 2) > 1e-6:
            raise ValueError('LocalLevel.level_scale_prior must come from build_bijector('
                             'build_inv_gamma_sd_prior(prior_level_sd)) or be left at its default.')
        return float(np.sqrt(pr.params['scale'] * 2 / kLocalLevelPriorSampleSize))

    @property
    def flags(self) -> int:
        """Layout flags of the wider model set (0 = the default model of build_default_model)."""
        from . import _native as nt
        f = 0
        if self._component(LocalLinearTrend) is not None:
            f |= nt.FLAG_LLT
        elif self._component(LocalLevel).level_scale_prior.family == 'LogNormal':
            f |= nt.FLAG_LEVEL_LOGNORMAL
        if self._component(LinearRegression) is not None:
            f |= nt.FLAG_DENSE_REG
        if self.observation_noise_scale_prior.family == 'LogNormal':
            f |= nt.FLAG_OBS_LOGNORMAL
        return f

    def _regression(self):
        return self._component(SparseLinearRegression) or self._component(LinearRegression)

    def param_names(self) -> List[str]:
        return [p.name for p in self.parameters]

    def param_slices(self):
        out, i = [], 0
        for p in self.parameters:
            n = int(np.prod(p.shape)) if p.shape else 1
            out.append((p.name, i, i + n, p.shape))
            i += n
        return out

    def engine(self, observed_time_series, Tf: int):
        """HBM-resident batch (N = 1) for this model; built once per (series, horizon)."""
        from .engine import DeviceBatch
        y = _response_array(observed_time_series)
        key = (y.tobytes(), int(Tf))
        if self._engine is not None and self._engine[0] == key:
            return self._engine[1]
        reg = self._regression()
        seas = [c for c in self.components if isinstance(c, Seasonal)]
        sea = seas[0] if seas else None
        sea2 = seas[1] if len(seas) > 1 else None
        X = None
        if reg is not None:
            X = reg.design_matrix.to_dense()
            if X.shape[0] < len(y) + Tf:
                raise ValueError(f'design_matrix has {X.shape[0]} rows; need pre + post = {len(y) + Tf}.')
            X = np.nan_to_num(X[:len(y) + Tf], nan=0.0)[None]
        obs_pr = self.observation_noise_scale_prior
        db = DeviceBatch(y[None, :], X, Tf=Tf, nseasons=sea.num_seasons if sea else 1,
                         season_duration=sea.num_steps_per_season if sea else 1,
                         prior_level_sd=self.prior_level_sd, flags=self.flags,
                         nseasons2=sea2.num_seasons if sea2 else 0,
                         season_duration2=sea2.num_steps_per_season if sea2 else 0)
        if obs_pr.family == 'SqrtInverseGamma':
            db.set_obs_prior_scale(obs_pr.params['scale'])
        self._engine = (key, db)
        return db


def _response_array(observed_time_series) -> np.ndarray:
    if isinstance(observed_time_series, (pd.DataFrame, pd.Series)):
        a = observed_time_series.values
    else:
        a = np.asarray(observed_time_series)
    return np.ascontiguousarray(a, dtype=np.float32).reshape(-1)


def check_input_model(model, pre_data: pd.DataFrame, post_data: pd.DataFrame) -> None:
    """Custom `model=`: only `Sum` models made of the supported components are accepted; anything
    else (e.g. a tfp.sts object) is rejected explicitly (north star; reference model.py:137-193)."""
    if not isinstance(model, Sum):
        raise ValueError('Input model must be of type StructuralTimeSeries.')
    for component in model.components:
        if isinstance(component, (SparseLinearRegression, LinearRegression)):
            expected = (len(pre_data) + len(post_data), len(pre_data.columns) - 1)
            if tuple(component.design_matrix.shape) != expected:
                raise ValueError(
                    'Customized Linear Regression Models must have total '
                    'points equal to pre_data and post_data points and '
                    'same number of covariates. Input design_matrix shape was '
                    f'{tuple(component.design_matrix.shape)} and expected '
                    f'{expected} '
                    'instead.')
            assert component.design_matrix.dtype == np.float32
        else:
            for parameter in component.parameters:
                assert parameter.prior.dtype == np.float32


def build_default_model(observed_time_series: pd.DataFrame, pre_data: pd.DataFrame, post_data: pd.DataFrame,
                        prior_level_sd: float, nseasons: int, season_duration: int) -> Sum:
    """LocalLevel [+ SparseLinearRegression if covariates] [+ Seasonal if nseasons > 1] with the
    reference's priors (model.py:239-321)."""
    obs_sd = observed_time_series.std(skipna=True, ddof=0).values[0]
    sd_prior = build_bijector(build_inv_gamma_sd_prior(prior_level_sd))
    obs_prior = build_bijector(build_inv_gamma_sd_prior(obs_sd / 2))
    components: List[Component] = [LocalLevel(level_scale_prior=sd_prior, observed_time_series=observed_time_series)]
    if len(pre_data.columns) > 1:
        complete = pd.concat([pre_data, post_data]).astype(np.float32).fillna(0)
        components.append(SparseLinearRegression(design_matrix=complete.iloc[:, 1:]))
    if nseasons > 1:
        components.append(Seasonal(num_seasons=nseasons, num_steps_per_season=season_duration,
                                   observed_time_series=observed_time_series))
    return Sum(components, observed_time_series=observed_time_series, observation_noise_scale_prior=obs_prior)


# ------------------------------------------------------------------------------------------------
# Fit
# ------------------------------------------------------------------------------------------------
def _split_params(model: Sum, theta: torch.Tensor):
    """theta [P, D] (device) -> list of CPU tensors [D, *shape] in `model.parameters` order."""
    th = theta.t().contiguous().cpu()
    return [(name, th[:, lo] if shape == () else th[:, lo:hi].reshape((-1,) + tuple(shape)))
            for name, lo, hi, shape in model.param_slices()]


def _horizon(model: Sum, observed_time_series) -> int:
    reg = model._regression()
    T = len(_response_array(observed_time_series))
    return max(reg.design_matrix.shape[0] - T, 0) if reg is not None else 0


def fit_model(model: Sum, observed_time_series: pd.DataFrame, method: str = 'hmc', seed: Optional[int] = None):
    """Posterior draws of the model parameters.

    'vi'  -> (dict name -> tensor[100, *shape], None)        model.py:366-386
    'hmc' -> (list of tensor[100, *shape], kernel_results)   model.py:356-365
    Both also stash the unconstrained draws on the model for the predictive path."""
    if method not in ('hmc', 'vi'):
        raise ValueError(f'Input method "{method}" not valid. Choose between "hmc" or "vi".')
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    db = model.engine(observed_time_series, max(_horizon(model, observed_time_series), model.__dict__.get('_Tf', 0)))
    if method == 'vi':
        mu, rho = db.vi_fit(1, 1, _VI_STEPS, 0.1, 0.01, seed)
        u = db.vi_draw(mu, rho, _VI_DRAWS, seed + 1)                       # [P, 100]
        model._u_draws = u
        return dict(_split_params(model, db.constrain(u))), None
    h = _HMC
    mu, rho = db.vi_fit(1, h['variational_sample_size'], h['num_variational_steps'], 0.1, 0.01, seed)
    u0 = db.vi_draw(mu, rho, 1, seed + 1)                                   # one q sample per chain
    step0 = torch.nn.functional.softplus(rho)                               # q stddev (unconstrained space)
    draws, stats = db.hmc_run(u0, step0, h['num_results'], h['num_warmup_steps'], None,
                              h['num_leapfrog_steps'], 0.75, seed + 2)
    u = draws[:, :, 0].t().contiguous()                                      # [P, 100]
    model._u_draws = u
    st = stats.cpu().numpy()
    kernel_results = {'accept_rate': float(st[0, 0]), 'warmup_accept_rate': float(st[1, 0]),
                      'step_size_factor': float(st[2, 0]), 'target_log_prob': float(st[3, 0]),
                      'initial_step_size': step0[:, 0].cpu().numpy()}
    return [t for _, t in _split_params(model, db.constrain(u))], kernel_results


def _unconstrained_from_samples(model: Sum, db, parameter_samples) -> torch.Tensor:
    """Accept the dict / list `fit_model` returned (constrained values) and map back to u [P, D]."""
    if isinstance(parameter_samples, dict):
        parts = [parameter_samples[name] for name in model.param_names()]
    else:
        parts = list(parameter_samples)
    cols = []
    for (name, lo, hi, shape), t in zip(model.param_slices(), parts):
        t = torch.as_tensor(np.asarray(t), dtype=torch.float32)
        cols.append(t.reshape(t.shape[0], -1))
    theta = torch.cat(cols, dim=1).t().contiguous().to(db.device)          # [P, D]
    return db.unconstrain(theta)


class _HostArray:
    """What `.sample()` / `.mean()` return: array-like with `.numpy()` (inferences.py:108-117)."""

    def __init__(self, t: torch.Tensor):
        self._t = t

    def numpy(self):
        return self._t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, item):
        return self.numpy()[item]

    @property
    def shape(self):
        return tuple(self._t.shape)


class OneStepDist:
    """Mixture over posterior draws of independent Normals per time step
    (tfp.sts.one_step_predictive, model.py:414-418).  Values are in MODEL units (standardised if
    the data were); the device path of inferences.py asks for original units directly."""

    def __init__(self, db, u):
        self.db, self.u = db, u
        self.pm, self.pv, self.fstate = db.filter_predict(u)
        self.event_shape = (db.T,)

    def sample_device(self, n, mu_sig=None, seed=None):
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else seed
        return self.db.onestep_sample(self.pm, self.pv, n, seed, mu_sig, want_mean=False)[0][0]   # [n, T]

    def mean_device(self, mu_sig=None):
        return self.db.onestep_sample(self.pm, self.pv, 0, 0, mu_sig, want_mean=True)[1][0]       # [T]

    def sample(self, n):
        return _HostArray(self.sample_device(n)[:, :, None])

    def mean(self):
        return _HostArray(self.mean_device())

    def stddev(self):
        m = self.pm.mean(dim=1)
        second = (self.pv + self.pm * self.pm).mean(dim=1)
        return _HostArray(torch.sqrt(torch.clamp(second - m * m, min=0.0)))


class PosteriorDist:
    """Mixture over posterior draws of forecast state-space models
    (tfp.sts.forecast, model.py:446-451): batch_shape [D], event_shape [Tf, 1]."""

    def __init__(self, db, u, fstate):
        self.db, self.u, self.fstate = db, u, fstate
        self.batch_shape = (u.shape[1],)
        self.event_shape = (db.Tf, 1)

    def sample_device(self, n, mu_sig=None, seed=None):
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else seed
        return self.db.forecast_sample(self.u, self.fstate, n, seed, mu_sig, want_mean=False)[0][0]  # [n, Tf]

    def mean_device(self, mu_sig=None):
        return self.db.forecast_sample(self.u, self.fstate, 0, 0, mu_sig, want_mean=True)[1][0]      # [Tf]

    def sample(self, n):
        return _HostArray(self.sample_device(n)[:, :, None])

    def mean(self):
        return _HostArray(self.mean_device()[:, None])


def build_one_step_dist(model: Sum, observed_time_series, parameter_samples) -> OneStepDist:
    db = model.engine(observed_time_series, max(_horizon(model, observed_time_series), model.__dict__.get('_Tf', 0)))
    u = _unconstrained_from_samples(model, db, parameter_samples)
    return OneStepDist(db, u)


def build_posterior_dist(model: Sum, observed_time_series, parameter_samples, num_steps_forecast: int
                         ) -> PosteriorDist:
    model._Tf = int(num_steps_forecast)
    db = model.engine(observed_time_series, int(num_steps_forecast))
    u = _unconstrained_from_samples(model, db, parameter_samples)
    _, _, fstate = db.filter_predict(u)
    return PosteriorDist(db, u, fstate)
