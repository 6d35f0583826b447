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
