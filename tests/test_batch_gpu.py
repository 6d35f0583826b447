"""Batched front door on the GPU: N series fitted / forecast / summarised in one pass; the known
injected effect must be recovered and agree with the single-series `CausalImpact` path."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(N, T, Tf, k, S, seed, effect):
    rng = np.random.default_rng(seed)
    Tt = T + Tf
    X = rng.normal(size=(N, Tt, k)).cumsum(axis=1) * 0.1 + rng.normal(size=(N, Tt, k)) * 0.5 + 5.0
    beta = rng.normal(size=(N, k)) + 1.0
    y = 20.0 + np.einsum('ntk,nk->nt', X, beta) + rng.normal(scale=0.5, size=(N, Tt))
    if S > 1:
        eff = rng.normal(scale=1.0, size=(N, S))
        eff -= eff.mean(axis=1, keepdims=True)
        y += eff[:, np.arange(Tt) % S]
    y[:, T:] += effect
    return y[:, :T].astype(np.float32), y[:, T:].astype(np.float32), X.astype(np.float32)


@pytest.mark.parametrize('method,S', [('vi', 1), ('vi', 7), ('hmc', 7)])
def test_batch_recovers_effect(method, S):
    from tfcausalimpact_b200.batch import causal_impact_batch
    N, T, Tf, k, effect = 24, 120, 28, 3, 3.0
    y_pre, y_post, X = _make(N, T, Tf, k, S, 5, effect)
    res = causal_impact_batch(y_pre, y_post, X, fit_method=method, nseasons=S, niter=1000, seed=11,
                              chains=2, num_results=60)
    summ = res['summary'].cpu().numpy()
    p = res['p_value'].cpu().numpy()
    assert np.all(np.isfinite(summ))
    abs_eff = summ[:, 4, 0]
    assert np.all(np.abs(abs_eff - effect) < 0.6), abs_eff
    assert np.all(summ[:, 2, 0] < summ[:, 1, 0]) and np.all(summ[:, 1, 0] < summ[:, 3, 0])
    assert np.all(p < 0.05)
    np.testing.assert_allclose(summ[:, 0, 0], y_post.mean(axis=1), rtol=1e-5)
    np.testing.assert_allclose(summ[:, 0, 1], y_post.sum(axis=1), rtol=1e-5)


def test_batch_agrees_with_single_series_api():
    from tfcausalimpact_b200 import CausalImpact
    from tfcausalimpact_b200.batch import causal_impact_batch
    N, T, Tf, k = 3, 100, 20, 2
    y_pre, y_post, X = _make(N, T, Tf, k, 1, 9, 2.0)
    res = causal_impact_batch(y_pre, y_post, X, fit_method='vi', niter=2000, seed=3)
    summ = res['summary'].cpu().numpy()
    for n in range(N):
        df = pd.DataFrame(np.concatenate([np.concatenate([y_pre[n], y_post[n]])[:, None], X[n]], axis=1))
        ci = CausalImpact(df, [0, T - 1], [T, T + Tf - 1], model_args={'niter': 2000, 'seed': 3})
        a, b = ci.summary_data.loc['predicted', 'average'], summ[n, 1, 0]
        sd = (ci.summary_data.loc['predicted_upper', 'average'] - ci.summary_data.loc['predicted_lower', 'average']) / 3.92
        # two independent 200-step single-sample VI fits differ by a sizeable fraction of the posterior sd
        assert abs(a - b) < 1.2 * sd + 1e-3, (a, b, sd)
        hw_a = ci.summary_data.loc['predicted_upper', 'average'] - ci.summary_data.loc['predicted_lower', 'average']
        hw_b = summ[n, 3, 0] - summ[n, 2, 0]
        assert 0.6 < hw_a / hw_b < 1.6


def test_causalimpact_batch_frontdoor():
    from tfcausalimpact_b200 import CausalImpact
    N, T, Tf, k = 8, 90, 21, 2
    y_pre, y_post, X = _make(N, T, Tf, k, 7, 21, 2.5)
    data = np.concatenate([np.concatenate([y_pre, y_post], axis=1)[:, :, None], X], axis=2)
    res = CausalImpact.batch(data, [0, T - 1], [T, T + Tf - 1], model_args={'nseasons': 7, 'niter': 500}, seed=4)
    assert len(res) == N
    df = res.to_frame()
    assert np.all(np.abs(df['abs_effect'].values - 2.5) < 0.8)
    assert res.summary_data(0).shape == (10, 2)
    assert res.summary(3).startswith('Posterior Inference {Causal Impact}')
    with pytest.raises(ValueError) as e:
        CausalImpact.batch(data, [0, T - 1], [T - 1, T + Tf - 1])
    assert 'post_period first value' in str(e.value) or 'Values in training data' in str(e.value)
    with pytest.raises(ValueError) as e:
        CausalImpact.batch(data, [0, T - 1], [T, T + Tf - 1], model_args={'fit_method': 'x'})
    assert str(e.value) == 'fit_method can be either "hmc" or "vi".'


def _frames_with_gap(N, T, Tf, k, seed, effect, drop_pre=17, drop_post=5):
    """N DataFrames on ONE daily DatetimeIndex with a missing day in each period (two of them on their own,
    different calendars so that several index groups exist)."""
    y_pre, y_post, X = _make(N, T, Tf, k, 1, seed, effect)
    full = pd.date_range('2021-03-01', periods=T + Tf, freq='D')
    keep = np.ones(T + Tf, dtype=bool)
    keep[[drop_pre, T + drop_post]] = False
    idx = pd.DatetimeIndex(full[keep].values)
    frames = []
    for n in range(N):
        vals = np.concatenate([np.concatenate([y_pre[n], y_post[n]])[:, None], X[n]], axis=1)[keep]
        frames.append(pd.DataFrame(vals, index=idx if n < N - 2 else pd.DatetimeIndex(idx.values),
                                   columns=['y'] + [f'x{j}' for j in range(k)]))
    return frames, full, keep


def test_batch_of_dataframes_with_gaps_matches_single_series_calls():
    """`CausalImpact.batch` on DataFrames with date periods and a calendar gap per period: same grids, masks,
    index labels and deterministic columns as N `CausalImpact(...)` calls; stochastic columns within MC error
    (/root/reference/causalimpact/data.py:31-165, inferences.py:208-246)."""
    from tfcausalimpact_b200 import CausalImpact
    N, T, Tf, k, effect = 6, 80, 20, 2, 2.0
    frames, full, keep = _frames_with_gap(N, T, Tf, k, 13, effect)
    pre_period = ['2021-03-01', full[T - 1].strftime('%Y-%m-%d')]
    post_period = [full[T].strftime('%Y-%m-%d'), full[T + Tf - 1].strftime('%Y-%m-%d')]
    res = CausalImpact.batch(frames, pre_period, post_period, model_args={'niter': 1000}, seed=2, inferences=True)
    assert len(res) == N
    for n in (0, N - 1):
        ci = CausalImpact(frames[n], pre_period, post_period, model_args={'niter': 1000, 'seed': 2})
        inf_b, inf_s = res.inferences_frame(n), ci.inferences
        assert list(inf_b.columns) == list(inf_s.columns)
        assert inf_b.index.equals(inf_s.index)
        assert len(inf_b) == T + Tf - 1                      # the masked post day is dropped, the pre gap is a row
        assert (inf_b.isna().values == inf_s.isna().values).all()
        np.testing.assert_allclose(inf_b['post_cum_y'].dropna().values, inf_s['post_cum_y'].dropna().values, rtol=1e-5)
        sd = (ci.summary_data.loc['predicted_upper', 'average'] - ci.summary_data.loc['predicted_lower', 'average']) / 3.92
        a, b = ci.summary_data.loc['predicted', 'average'], res.summary_data(n).loc['predicted', 'average']
        assert abs(a - b) < 1.5 * sd + 1e-3
        np.testing.assert_allclose(res.summary_data(n).loc['actual'].values, ci.summary_data.loc['actual'].values,
                                   rtol=1e-5)
        m = inf_s['complete_preds_means'].notna().values
        assert np.abs(inf_b['complete_preds_means'].values[m] - inf_s['complete_preds_means'].values[m]).mean() < 1.0 * sd + 0.05
        assert (inf_b['complete_preds_lower'].values[m] <= inf_b['complete_preds_upper'].values[m]).all()
    assert np.all(np.abs(res.to_frame()['abs_effect'].values - effect) < 0.8)


def test_batch_long_format_and_constant_covariate():
    """Long-format input (series id + date columns) takes the reshape path; a covariate that is constant over the
    pre-period (0/0 in the standardisation) no longer poisons the series."""
    from tfcausalimpact_b200 import CausalImpact
    N, T, Tf, k = 5, 60, 14, 2
    y_pre, y_post, X = _make(N, T, Tf, k, 1, 31, 1.5)
    X[:, :T, 1] = 0.0
    X[:, T:, 1] = 1.0                                            # intervention dummy
    dates = pd.date_range('2022-01-01', periods=T + Tf, freq='D')
    rows = []
    for n in range(N):
        rows.append(pd.DataFrame({'id': f's{n}', 'date': dates, 'y': np.concatenate([y_pre[n], y_post[n]]),
                                  'x0': X[n, :, 0], 'x1': X[n, :, 1]}))
    long = pd.concat(rows, ignore_index=True)
    res = CausalImpact.batch(long, ['2022-01-01', dates[T - 1]], [dates[T], dates[-1]], model_args={'niter': 400},
                             series_col='id', time_col='date', seed=1)
    assert len(res) == N and list(res.labels) == [f's{n}' for n in range(N)]
    fr = res.to_frame()
    assert np.isfinite(fr.values).all()
    shuffled = long.sample(frac=1.0, random_state=0)
    res2 = CausalImpact.batch(shuffled, ['2022-01-01', dates[T - 1]], [dates[T], dates[-1]], model_args={'niter': 400},
                              series_col='id', time_col='date', seed=1)
    order = [list(res2.labels).index(f's{n}') for n in range(N)]
    np.testing.assert_allclose(res2.to_frame()['actual'].values[order], fr['actual'].values, rtol=1e-6)
