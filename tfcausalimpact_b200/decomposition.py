"""Component decomposition of a fitted model (the counterpart of `tfp.sts.decompose_by_component`
/ `tfp.sts.decompose_forecast_by_component`, which the reference's notebook calls on `ci.model`,
`ci.observed_time_series`, `ci.model_samples` -- notebooks/getting_started.ipynb:125,144).

Both return an OrderedDict keyed by the model's component objects (use `component.name`); every
value exposes `.mean()` and `.stddev()` (array-likes with `.numpy()`), in MODEL units (standardised
when the data were), like the TFP distributions they replace."""
from collections import OrderedDict

from . import model as cimodel


class ComponentDist:
    def __init__(self, mean, sd):
        self._mean, self._sd = mean, sd

    def mean(self):
        return cimodel._HostArray(self._mean)

    def stddev(self):
        return cimodel._HostArray(self._sd)


def _components(model, mean, sd):
    out = OrderedDict()
    seasonal_row = 2
    for comp in model.components:
        if isinstance(comp, cimodel.Seasonal):
            row = seasonal_row          # first Seasonal -> row 2, second -> row 3
            seasonal_row += 1
        else:
            row = {cimodel.LocalLevel: 0, cimodel.LocalLinearTrend: 0, cimodel.SparseLinearRegression: 1,
                   cimodel.LinearRegression: 1}[type(comp)]
        out[comp] = ComponentDist(mean[row, 0], sd[row, 0])
    return out


def decompose_by_component(model, observed_time_series, parameter_samples):
    """Smoothed marginal of every component over the observed (pre-intervention) period."""
    db = model.engine(observed_time_series, max(cimodel._horizon(model, observed_time_series),
                                                model.__dict__.get('_Tf', 0)))
    u = cimodel._unconstrained_from_samples(model, db, parameter_samples)
    mean, sd = db.decompose(u)
    return _components(model, mean, sd)


def decompose_forecast_by_component(model, forecast_dist, parameter_samples):
    """Forecast marginal of every component over the post-intervention horizon; `forecast_dist` is
    the `PosteriorDist` returned by `build_posterior_dist` (`ci.posterior_dist`)."""
    db = forecast_dist.db
    u = cimodel._unconstrained_from_samples(model, db, parameter_samples)
    mean, sd = db.decompose(u, forecast_dist.fstate)
    return _components(model, mean, sd)
