"""The reference's own deterministic reduction test, restated against the GPU reductions:
duck-typed distributions exposing `.sample(n)` / `.mean()` exactly like the fakes of
/root/reference/tests/test_inferences.py:47-345 (the de-facto plugin protocol), every one of the 16
columns recomputed independently with NumPy, plus summarize / p-value (:348-455)."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


class _Arr:
    def __init__(self, a):
        self._a = np.asarray(a, dtype=np.float32)

    def numpy(self):
        return self._a


def _fakes(pre_len, post_len, one_step_mean=3.0, posterior_mean=7.5):
    class OneStepDist:
        def sample(self, niter):
            return _Arr((np.tile(np.arange(3.1, 3.1 + pre_len, 1.0), (niter, 1)) +
                         np.arange(niter).reshape(-1, 1))[..., None])

        def mean(self):
            return np.ones((pre_len, 1)) * one_step_mean

    class PosteriorDist:
        def sample(self, niter):
            return _Arr((np.tile(np.arange(7.1, 7.1 + post_len, 1.0), (niter, 1)) +
                         np.arange(niter).reshape(-1, 1))[..., None])

        def mean(self):
            return np.ones((post_len, 1)) * posterior_mean
    return OneStepDist(), PosteriorDist()


def test_compile_posterior_inferences_reference_fakes():
    from tfcausalimpact_b200 import inferences as inferrer
    data = pd.DataFrame(np.arange(10, dtype=np.float32))
    pre_data, post_data = data.iloc[:3], data.iloc[7:]
    alpha, mu, sig, niter = 0.05, 1.0, 2.0, 10
    mask = np.array([True] * len(post_data))
    one, post = _fakes(3, 3)
    inf = inferrer.compile_posterior_inferences(data.index, mask, pre_data, post_data, one, post, (mu, sig),
                                                alpha=alpha, niter=niter)
    assert isinstance(inf.index, pd.RangeIndex) and len(inf) == 6
    assert list(inf.columns) == inferrer.COLUMNS
    # expectations exactly as the reference computes them
    np.testing.assert_allclose(inf['complete_preds_means'].values, [7, 7, 7, 16, 16, 16], rtol=1e-6)
    pre_s = one.sample(niter).numpy()[..., 0] * sig + mu
    post_s = post.sample(niter).numpy()[..., 0] * sig + mu
    lo, hi = np.percentile(pre_s, [2.5, 97.5], axis=0)
    plo, phi = np.percentile(post_s, [2.5, 97.5], axis=0)
    np.testing.assert_allclose(inf['complete_preds_lower'].values, np.concatenate([lo, plo]), rtol=1e-5)
    np.testing.assert_allclose(inf['complete_preds_upper'].values, np.concatenate([hi, phi]), rtol=1e-5)
    np.testing.assert_allclose(inf['post_preds_means'].values[3:], [16, 16, 16], rtol=1e-6)
    assert inf['post_preds_means'].iloc[:3].isna().all()
    np.testing.assert_allclose(inf['post_preds_lower'].values[3:], plo, rtol=1e-5)
    np.testing.assert_allclose(inf['post_preds_upper'].values[3:], phi, rtol=1e-5)
    y_post = post_data.iloc[:, 0].values
    np.testing.assert_allclose(inf['post_cum_y'].values[2:], np.concatenate([[0], np.cumsum(y_post)]), rtol=1e-6)
    assert inf['post_cum_y'].iloc[:2].isna().all()
    np.testing.assert_allclose(inf['post_cum_preds_means'].values[2:], [0, 16, 32, 48], rtol=1e-6)
    clo, chi = np.percentile(np.cumsum(post_s, axis=1), [2.5, 97.5], axis=0)
    np.testing.assert_allclose(inf['post_cum_preds_lower'].values[2:], np.concatenate([[0], clo]), rtol=1e-5)
    np.testing.assert_allclose(inf['post_cum_preds_upper'].values[2:], np.concatenate([[0], chi]), rtol=1e-5)
    y_all = np.concatenate([pre_data.iloc[:, 0].values, y_post])
    np.testing.assert_allclose(inf['point_effects_means'].values, y_all - inf['complete_preds_means'].values, rtol=1e-6)
    np.testing.assert_allclose(inf['point_effects_lower'].values, y_all - inf['complete_preds_upper'].values, rtol=1e-5,
                               atol=1e-5)
    np.testing.assert_allclose(inf['point_effects_upper'].values, y_all - inf['complete_preds_lower'].values, rtol=1e-5,
                               atol=1e-5)
    np.testing.assert_allclose(inf['post_cum_effects_means'].values[2:],
                               np.concatenate([[0], np.cumsum(y_post - 16.0)]), rtol=1e-6)
    elo, ehi = np.percentile(np.cumsum(y_post - post_s, axis=1), [2.5, 97.5], axis=0)
    np.testing.assert_allclose(inf['post_cum_effects_lower'].values[2:], np.concatenate([[0], elo]), rtol=1e-5)
    np.testing.assert_allclose(inf['post_cum_effects_upper'].values[2:], np.concatenate([[0], ehi]), rtol=1e-5)


def test_summarize_and_p_value_reference_semantics():
    from tfcausalimpact_b200 import inferences as inferrer
    post_preds_means = pd.Series([7.0, 8.0, 9.0])
    post_data = pd.DataFrame([7.5, 8.5, 9.5], dtype=np.float32)
    sims = (np.tile(np.arange(7.1, 10.1, 1.0), (10, 1)) + np.arange(10).reshape(-1, 1)).astype(np.float32)
    s = inferrer.summarize_posterior_inferences(post_preds_means, post_data, sims, 0.05)
    mlo, mhi = np.percentile(sims.mean(axis=1), [2.5, 97.5])
    slo, shi = np.percentile(sims.sum(axis=1), [2.5, 97.5])
    assert list(s.index) == inferrer.SUMMARY_INDEX and list(s.columns) == ['average', 'cumulative']
    np.testing.assert_allclose(s.loc['actual'].values, [8.5, 25.5], rtol=1e-6)
    np.testing.assert_allclose(s.loc['predicted'].values, [8.0, 24.0], rtol=1e-6)
    np.testing.assert_allclose(s.loc['predicted_lower'].values, [mlo, slo], rtol=1e-5)
    np.testing.assert_allclose(s.loc['predicted_upper'].values, [mhi, shi], rtol=1e-5)
    np.testing.assert_allclose(s.loc['abs_effect'].values, [0.5, 1.5], rtol=1e-5)
    np.testing.assert_allclose(s.loc['abs_effect_lower'].values, [8.5 - mhi, 25.5 - shi], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(s.loc['rel_effect'].values, [0.5 / 8, 1.5 / 24], rtol=1e-5)
    np.testing.assert_allclose(s.loc['rel_effect_upper'].values, [(8.5 - mlo) / 8, (25.5 - slo) / 24], rtol=1e-4,
                               atol=1e-6)
    # reference tests/test_inferences.py:448-455: p == 2 / 11
    simulated_ys = (np.tile(np.arange(7, 10, 1), (10, 1)) + np.arange(10).reshape(-1, 1)).astype(np.float32)
    assert abs(inferrer.compute_p_value(simulated_ys, 46) - 2 / 11) < 1e-7


def test_rhat_across_chains_cfg3_shape():
    """SURVEY 8(d) gate: split R-hat across 8 chains on a configs[2]-shaped sample."""
    import torch
    from helpers import synth_problem
    from tfcausalimpact_b200.engine import DeviceBatch
    N, C, T, Tf, k, S = 16, 8, 365, 30, 10, 7
    y, _, X = synth_problem(N, T, Tf, k, S, 1, seed=4)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    mu, rho = db.vi_fit(C, 5, 150, 0.1, 0.01, seed=1)
    u = db.vi_draw(mu, rho, 1, seed=2)
    step0 = torch.nn.functional.softplus(rho).contiguous()
    draws, stats = db.hmc_run(u, step0, 600, 300, 240, 15, 0.75, seed=3)       # [R, P, N*C]
    R, P, B = draws.shape
    d = draws.view(R, P, N, C).double()
    def split_rhat(x):                                                           # x [R, ..., N, C] -> [..., N]
        half = x.shape[0] // 2
        ch = torch.cat([x[:half], x[half:2 * half]], dim=-1)                     # split chains: 2C per series
        m = ch.mean(dim=0)
        W = ch.var(dim=0, unbiased=True).mean(dim=-1)
        Bv = half * m.var(dim=-1, unbiased=True)
        return torch.sqrt(((half - 1) / half * W + Bv / half) / W)
    d = draws.view(R, P, N, C).double()
    rh_scales = split_rhat(d)[[0, 1, P - 1]]                                     # sigma_obs, sigma_level, sigma_drift
    assert float(stats[0].mean()) > 0.5
    assert float(rh_scales.median()) < 1.05, rh_scales.median()
    assert float((rh_scales < 1.2).double().mean()) > 0.9
    # the regression block, where HMC actually struggles (horseshoe funnel): realised weights
    # beta = wn * (lsn sqrt(lsv)) * (gsn sqrt(gsv) 0.1) and the log global / local scales, from the CONSTRAINED draws
    theta = torch.stack([db.constrain(draws[r].contiguous()) for r in range(R)]).view(R, P, N, C).double()
    gsv, gsn = theta[:, 2], theta[:, 3]
    lsv, lsn, wn = theta[:, 4:4 + k], theta[:, 4 + k:4 + 2 * k], theta[:, 4 + 2 * k:4 + 3 * k]
    tau = gsn * gsv.sqrt() * 0.1                                                 # [R, N, C]
    lam = lsn * lsv.sqrt()                                                       # [R, k, N, C]
    beta = wn * lam * tau[:, None]
    rh_beta, rh_tau, rh_lam = split_rhat(beta), split_rhat(tau.log()), split_rhat(lam.log())
    msg = (float(rh_beta.median()), float(rh_beta.max()), float(rh_tau.median()), float(rh_tau.max()),
           float(rh_lam.median()), float(rh_lam.quantile(0.95)))
    print('split R-hat (median, max): beta %.3f %.3f | log tau %.3f %.3f | log lambda (median, q95) %.3f %.3f' % msg)
    # what users consume -- the realised weights -- must mix; the horseshoe's global scale sits at the neck of the
    # funnel and mixes slowly under the reference's sampler settings (identity mass, L = 15, 300 warm-up steps: the
    # same algorithm), so its gate only guards against divergence, the measured values are printed
    assert float(rh_beta.median()) < 1.05 and float((rh_beta < 1.2).double().mean()) > 0.95, msg
    assert float(rh_lam.median()) < 1.1 and float((rh_lam < 1.5).double().mean()) > 0.9, msg
    assert float(rh_tau.median()) < 2.0, msg
