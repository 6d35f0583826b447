"""Component decomposition (ci_decompose_by_component / ci_decompose_forecast) against the oracle's
dense Rauch-Tung-Striebel smoother / moment propagation.  fp32 kernel vs fp64 oracle; the smoother
solves a d x d system per step, so tolerances are rtol 2e-3 on means (atol 2e-3 of the series
scale) and rtol 5e-3 on standard deviations."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from helpers import O, synth_problem

pytestmark = pytest.mark.gpu

CASES = [(2, 5, 60, 8, 2, 7, 1, 0.0), (3, 4, 50, 6, 0, 1, 1, 0.1), (2, 3, 48, 9, 3, 4, 2, 0.1), (1, 6, 40, 5, 1, 12, 1, 0.0)]


@pytest.mark.parametrize('N,G,T,Tf,k,S,r,nan_frac', CASES)
def test_decomposition_matches_oracle(N, G, T, Tf, k, S, r, nan_frac):
    from tfcausalimpact_b200.engine import DeviceBatch
    L = O.Layout(T, Tf, k, S, r)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=41, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r)
    B = N * G
    rng = np.random.default_rng(3)
    u = (rng.normal(size=(L.P, B)) * 0.4).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 1.5
    ud = torch.as_tensor(u).cuda()
    series = torch.arange(B) // G
    uo = torch.as_tensor(u.T.astype(np.float64))
    rows = {'LocalLevel': 0, 'SparseLinearRegression': 1, 'Seasonal': 2}
    mean, sd = db.decompose(ud)
    ref = O.smooth_components(prob, uo, series, G)
    for name, (m_o, s_o) in ref.items():
        np.testing.assert_allclose(mean[rows[name]].cpu().numpy(), m_o.numpy(), rtol=2e-3, atol=2e-3, err_msg=name)
        np.testing.assert_allclose(sd[rows[name]].cpu().numpy(), s_o.numpy(), rtol=5e-3, atol=1e-3, err_msg=name)
    _, _, fs = db.filter_predict(ud)
    meanf, sdf = db.decompose(ud, fs)
    reff = O.forecast_components(prob, uo, series, G)
    for name, (m_o, s_o) in reff.items():
        np.testing.assert_allclose(meanf[rows[name]].cpu().numpy(), m_o.numpy(), rtol=1e-3, atol=1e-3, err_msg=name)
        np.testing.assert_allclose(sdf[rows[name]].cpu().numpy(), s_o.numpy(), rtol=2e-3, atol=1e-4, err_msg=name)


# layouts on the warp-per-lane kernels (Durbin-Koopman backward recursion, ci_bigd.cu): large nseasons, the
# wider model set of custom `model=` and two Seasonal components
BIG_CASES = [
    # N, G, T, Tf, k, S, r, nan_frac, flags, S2, r2
    (2, 3, 60, 9, 2, 9, 1, 0.0, 0, 1, 1),
    (1, 3, 80, 20, 1, 52, 1, 0.1, 0, 1, 1),
    (2, 2, 50, 8, 3, 1, 1, 0.0, O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 1, 1),
    (1, 3, 70, 10, 0, 7, 1, 0.05, O.FLAG_LLT | O.FLAG_OBS_LOGNORMAL, 1, 1),
    (1, 2, 120, 12, 3, 4, 1, 0.0, O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 52, 1),
    (2, 2, 60, 14, 2, 7, 1, 0.1, 0, 4, 3),
]


@pytest.mark.parametrize('N,G,T,Tf,k,S,r,nan_frac,flags,S2,r2', BIG_CASES)
def test_decomposition_large_latent_and_wider_models(N, G, T, Tf, k, S, r, nan_frac, flags, S2, r2):
    from tfcausalimpact_b200.engine import DeviceBatch
    L = O.Layout(T, Tf, k, S, r, flags, S2, r2)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=43, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r, flags=flags,
                     nseasons2=S2 if S2 > 1 else 0, season_duration2=r2)
    B = N * G
    rng = np.random.default_rng(3)
    u = (rng.normal(size=(L.P, B)) * 0.4).astype(np.float32)
    u[0] -= 1.0
    u[1:1 + L.nt] -= 1.5
    ud = torch.as_tensor(u).cuda()
    series = torch.arange(B) // G
    uo = torch.as_tensor(u.T.astype(np.float64))
    rows = {'LocalLevel': 0, 'SparseLinearRegression': 1, 'Seasonal': 2, 'Seasonal2': 3}
    mean, sd = db.decompose(ud)
    assert mean.shape[0] == (4 if S2 > 1 else 3)
    ref = O.smooth_components(prob, uo, series, G)
    assert ('Seasonal2' in ref) == (S2 > 1)
    for name, (m_o, s_o) in ref.items():
        np.testing.assert_allclose(mean[rows[name]].cpu().numpy(), m_o.numpy(), rtol=2e-3, atol=2e-3, err_msg=name)
        np.testing.assert_allclose(sd[rows[name]].cpu().numpy(), s_o.numpy(), rtol=5e-3, atol=1e-3, err_msg=name)
    _, _, fs = db.filter_predict(ud)
    meanf, sdf = db.decompose(ud, fs)
    reff = O.forecast_components(prob, uo, series, G)
    for name, (m_o, s_o) in reff.items():
        np.testing.assert_allclose(meanf[rows[name]].cpu().numpy(), m_o.numpy(), rtol=1e-3, atol=1e-3, err_msg=name)
        np.testing.assert_allclose(sdf[rows[name]].cpu().numpy(), s_o.numpy(), rtol=2e-3, atol=1e-4, err_msg=name)


def test_decompose_front_door():
    from tfcausalimpact_b200 import CausalImpact
    from tfcausalimpact_b200.decomposition import decompose_by_component, decompose_forecast_by_component
    rng = np.random.default_rng(0)
    n = 150
    x = rng.normal(size=n).cumsum() * 0.1 + 10
    seas = np.tile([1.0, -0.5, 0.2, -0.7, 0.0, 0.6, -0.6], n // 7 + 1)[:n]
    y = 1.5 * x + seas + rng.normal(scale=0.2, size=n)
    df = pd.DataFrame({'y': y, 'x': x})
    np.random.seed(2)
    ci = CausalImpact(df, [0, 119], [120, 149], model_args={'nseasons': 7})
    comps = decompose_by_component(ci.model, ci.observed_time_series, ci.model_samples)
    names = [c.name for c in comps]
    assert names == ['LocalLevel', 'SparseLinearRegression', 'Seasonal']
    total = sum(d.mean().numpy() for d in comps.values())
    yn = ci.normed_pre_data.iloc[:, 0].values
    # the smoothed components add up to the (standardised) series up to the observation noise
    assert np.abs(total + ci.observed_time_series.iloc[:, 0].mean() - yn).std() < 0.25
    seas_hat = list(comps.values())[2].mean().numpy()
    assert np.corrcoef(seas_hat[14:], seas[14:120])[0, 1] > 0.9
    fc = decompose_forecast_by_component(ci.model, ci.posterior_dist, ci.model_samples)
    assert [c.name for c in fc] == names
    for d in fc.values():
        assert d.mean().numpy().shape == (30,) and np.all(d.stddev().numpy() >= 0)
    assert list(fc.values())[0].stddev().numpy()[-1] > list(fc.values())[0].stddev().numpy()[0]


def test_decompose_front_door_custom_model_two_seasonals():
    """Decomposition of a custom `model=` with a LocalLinearTrend, dense regression and two Seasonal
    components (the notebook's cell-116 model shape) through the public helpers."""
    from tfcausalimpact_b200 import CausalImpact
    from tfcausalimpact_b200 import model as M
    from tfcausalimpact_b200.decomposition import decompose_by_component, decompose_forecast_by_component
    from tfcausalimpact_b200.misc import standardize
    rng = np.random.default_rng(7)
    n, npre = 132, 120
    x = rng.normal(size=(n, 2)).cumsum(axis=0) * 0.1 + 5
    s4 = np.array([0.8, -0.3, -0.6, 0.1])[np.arange(n) % 4]
    s12 = 1.2 * np.sin(2 * np.pi * np.arange(n) / 12.0)
    y = 3 + 0.01 * np.arange(n) + 1.0 * x[:, 0] + s4 + s12 + rng.normal(scale=0.15, size=n)
    data = pd.DataFrame(np.column_stack([y, x]), columns=['y', 'x0', 'x1']).astype(np.float32)
    normed = standardize(data.iloc[:npre])[0]
    mu, sig = data.iloc[:npre].mean(), data.iloc[:npre].std(ddof=0)
    normed_all = ((data - mu) / sig).astype(np.float32)
    obs = normed_all['y'].iloc[:npre]
    model = M.Sum([M.LocalLinearTrend(observed_time_series=obs), M.LinearRegression(design_matrix=normed_all.iloc[:, 1:]),
                   M.Seasonal(num_seasons=4, observed_time_series=obs, name='Quarter'),
                   M.Seasonal(num_seasons=12, observed_time_series=obs, name='Year')], observed_time_series=obs)
    np.random.seed(8)
    ci = CausalImpact(data, [0, npre - 1], [npre, n - 1], model=model)
    comps = decompose_by_component(ci.model, ci.observed_time_series, ci.model_samples)
    assert [c.name for c in comps] == ['LocalLinearTrend', 'LinearRegression', 'Quarter', 'Year']
    vals = [d.mean().numpy() for d in comps.values()]
    assert all(v.shape == (npre,) and np.isfinite(v).all() for v in vals)
    assert np.corrcoef(vals[2][24:], s4[24:npre])[0, 1] > 0.8
    assert np.corrcoef(vals[3][24:], s12[24:npre])[0, 1] > 0.8
    fc = decompose_forecast_by_component(ci.model, ci.posterior_dist, ci.model_samples)
    assert [c.name for c in fc] == [c.name for c in comps]
    for d in fc.values():
        assert d.mean().numpy().shape == (n - npre,) and np.all(d.stddev().numpy() >= 0)
