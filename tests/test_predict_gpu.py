"""GPU parity of the predictive path (ci_filter_predict / ci_onestep_sample / ci_forecast_sample)
against the oracle's restatement of tfp.sts.one_step_predictive / tfp.sts.forecast.
Filter outputs: rtol 1e-4 (north_star).  Samplers: RNG parity with TF is impossible, so they are
checked against the oracle's closed-form forecast moments within Monte-Carlo error."""
import numpy as np
import pytest
import torch

from helpers import O, synth_problem

pytestmark = pytest.mark.gpu

CASES = [(3, 4, 40, 6, 2, 4, 1, 0.0), (2, 5, 30, 8, 0, 1, 1, 0.1), (2, 3, 45, 9, 1, 3, 2, 0.1),
         (2, 2, 60, 14, 3, 7, 1, 0.0)]


def _setup(N, T, Tf, k, S, r, seed, nan_frac):
    from tfcausalimpact_b200.engine import DeviceBatch
    L = O.Layout(T, Tf, k, S, r)
    y, ypost, X = synth_problem(N, T, Tf, k, S, r, seed=seed, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r)
    return L, prob, db


def _draws(L, B, seed=1):
    rng = np.random.default_rng(seed)
    u = (rng.normal(size=(L.P, B)) * 0.4).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    return u


@pytest.mark.parametrize('N,G,T,Tf,k,S,r,nan_frac', CASES)
def test_filter_predict_matches_oracle(N, G, T, Tf, k, S, r, nan_frac):
    L, prob, db = _setup(N, T, Tf, k, S, r, 21, nan_frac)
    B = N * G
    u = _draws(L, B)
    pm, pv, fs = db.filter_predict(torch.as_tensor(u).cuda())
    series = torch.arange(B) // G
    mean_o, var_o, m_o, P_o, _, _ = O.one_step_predictive(prob, torch.as_tensor(u.T.astype(np.float64)), series)
    np.testing.assert_allclose(pm.cpu().numpy().T, mean_o.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pv.cpu().numpy().T, var_o.numpy(), rtol=1e-4)
    fs = fs.cpu().numpy()
    d = L.d
    np.testing.assert_allclose(fs[:d].T, m_o.numpy(), rtol=1e-4, atol=1e-5)
    # packed upper covariance + its Cholesky factor
    iu = np.triu_indices(d)
    Pd = np.zeros((B, d, d))
    Pd[:, iu[0], iu[1]] = fs[d:d + len(iu[0])].T
    Pd = Pd + np.triu(Pd, 1).transpose(0, 2, 1)
    scale = np.abs(P_o.numpy()).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(Pd - P_o.numpy()) <= 1e-4 * scale)
    il = np.tril_indices(d)
    Lf = np.zeros((B, d, d))
    Lf[:, il[0], il[1]] = fs[d + len(iu[0]):].T
    assert np.all(np.abs(Lf @ Lf.transpose(0, 2, 1) - Pd) <= 2e-4 * scale)


@pytest.mark.parametrize('N,G,T,Tf,k,S,r,nan_frac', CASES[:3])
def test_forecast_mean_and_sampler_moments(N, G, T, Tf, k, S, r, nan_frac):
    L, prob, db = _setup(N, T, Tf, k, S, r, 22, nan_frac)
    B = N * G
    u = _draws(L, B, 3)
    ud = torch.as_tensor(u).cuda()
    pm, pv, fs = db.filter_predict(ud)
    nsamp = 4000
    mu_sig = torch.tensor([[10.0] * N, [2.0] * N]).cuda()
    sims, mean = db.forecast_sample(ud, fs, nsamp, seed=5, mu_sig=mu_sig)
    series = torch.arange(B) // G
    means_o, vars_o = O.forecast_moments(prob, torch.as_tensor(u.T.astype(np.float64)), series)
    means_o = means_o.numpy().reshape(N, G, Tf) * 2.0 + 10.0
    vars_o = vars_o.numpy().reshape(N, G, Tf) * 4.0
    # .mean(): deterministic, mixture mean over draws
    np.testing.assert_allclose(mean.cpu().numpy(), means_o.mean(axis=1), rtol=1e-4, atol=1e-4)
    # .sample(): mixture moments within MC error (5 sigma)
    s = sims.cpu().numpy().astype(np.float64)                     # [N, nsamp, Tf]
    mix_mean = means_o.mean(axis=1)
    mix_var = (vars_o + means_o ** 2).mean(axis=1) - mix_mean ** 2
    se = np.sqrt(mix_var / nsamp)
    assert np.all(np.abs(s.mean(axis=1) - mix_mean) < 5 * se)
    assert np.all(np.abs(s.var(axis=1) / mix_var - 1.0) < 0.25)
    # trajectories are joint: increments are positively correlated through the level
    assert np.all(np.isfinite(s))


def test_onestep_sampler_moments_and_mean():
    N, G, T, Tf, k, S = 2, 6, 50, 5, 2, 4
    L, prob, db = _setup(N, T, Tf, k, S, 1, 23, 0.0)
    B = N * G
    u = _draws(L, B, 4)
    pm, pv, fs = db.filter_predict(torch.as_tensor(u).cuda())
    nsamp = 4000
    sims, mean = db.onestep_sample(pm, pv, nsamp, seed=8)
    pmn = pm.cpu().numpy().reshape(T, N, G).astype(np.float64)
    pvn = pv.cpu().numpy().reshape(T, N, G).astype(np.float64)
    np.testing.assert_allclose(mean.cpu().numpy(), pmn.mean(axis=2).T, rtol=1e-5, atol=1e-6)
    mix_mean = pmn.mean(axis=2).T
    mix_var = (pvn + pmn ** 2).mean(axis=2).T - mix_mean ** 2
    s = sims.cpu().numpy().astype(np.float64)
    assert np.all(np.abs(s.mean(axis=1) - mix_mean) < 5 * np.sqrt(mix_var / nsamp))
    assert np.all(np.abs(s.var(axis=1) / mix_var - 1.0) < 0.25)
    # every draw of the mixture is used about equally often: chi-square-free sanity via the seed
    sims2, _ = db.onestep_sample(pm, pv, nsamp, seed=9, want_mean=False)
    assert float((sims2 - sims).abs().max()) > 0.0


@pytest.mark.parametrize('tiled', [1, 0])
@pytest.mark.parametrize('N,G,T,nsamp', [(3, 5, 37, 4099), (2, 7, 31, 33), (1, 100, 365, 1000), (5, 1, 70, 258)])
def test_onestep_sampler_edge_shapes(tiled, N, G, T, nsamp):
    """Both sampler kernels (tiled lane-per-step with aligned runs / sample-per-thread gather) on ragged shapes:
    several slabs per series, a partial last quad, T below one block, one draw; every element written, per-step
    moments of the mixture, no correlation between consecutive steps or consecutive samples."""
    import ctypes
    from tfcausalimpact_b200 import _native as nt
    Tf, k, S = 5, 1, 4
    L, prob, db = _setup(N, T, Tf, k, S, 1, 29, 0.0)
    u = _draws(L, N * G, 5)
    pm, pv, fs = db.filter_predict(torch.as_tensor(u).cuda())
    nt.check(nt.lib().ci_set_option(b'onestep_tiled', ctypes.c_long(tiled)), 'opt')
    try:
        # NaN-prefill the block the caching allocator hands out next: an element the kernel skips stays NaN
        x = torch.full((N, nsamp, T), float('nan'), device='cuda')
        del x
        sims, mean = db.onestep_sample(pm, pv, nsamp, seed=11)
    finally:
        nt.check(nt.lib().ci_set_option(b'onestep_tiled', ctypes.c_long(1)), 'opt')
    s = sims.cpu().numpy().astype(np.float64)
    assert s.shape == (N, nsamp, T) and np.all(np.isfinite(s))
    pmn = pm.cpu().numpy().reshape(T, N, G).astype(np.float64)
    pvn = pv.cpu().numpy().reshape(T, N, G).astype(np.float64)
    mix_mean = pmn.mean(axis=2).T
    np.testing.assert_allclose(mean.cpu().numpy(), mix_mean, rtol=1e-5, atol=1e-6)
    mix_var = (pvn + pmn ** 2).mean(axis=2).T - mix_mean ** 2
    assert np.all(np.abs(s.mean(axis=1) - mix_mean) < 5 * np.sqrt(mix_var / nsamp))
    if nsamp >= 1000:
        assert np.all(np.abs(s.var(axis=1) / mix_var - 1.0) < 0.3)
    if G == 1:
        # one draw: y_t independent over t and over samples -> standardised residuals are white noise
        z = (s - mix_mean[:, None, :]) / np.sqrt(mix_var)[:, None, :]
        n_el = z[:, :, 1:].size
        assert abs((z[:, :, 1:] * z[:, :, :-1]).mean()) < 5 / np.sqrt(n_el)
        assert abs((z[:, 1:, :] * z[:, :-1, :]).mean()) < 5 / np.sqrt(n_el)


def test_forecast_sampler_single_draw_covariance():
    """With a single posterior draw the samples are one Gaussian process: check the variance of the
    path SUM (needs the cross-time covariance, i.e. joint trajectories) against a dense oracle."""
    N, G, T, Tf, k, S = 1, 1, 40, 12, 0, 4
    L, prob, db = _setup(N, T, Tf, k, S, 1, 24, 0.0)
    u = _draws(L, 1, 6)
    ud = torch.as_tensor(u).cuda()
    pm, pv, fs = db.filter_predict(ud)
    nsamp = 8000
    sims, _ = db.forecast_sample(ud, fs, nsamp, seed=1)
    s = sims.cpu().numpy()[0].astype(np.float64)
    # dense covariance of (y_T .. y_T+Tf-1) from the oracle's matrices
    series = torch.zeros(1, dtype=torch.long)
    _, _, m, P, th, ssm = O.one_step_predictive(prob, torch.as_tensor(u.T.astype(np.float64)), series)
    d = L.d
    H = ssm.H.numpy()
    P = P.numpy()[0]
    # propagate state covariance and cross-covariances
    Fs = [ssm.F(L.T + t).numpy() for t in range(Tf)]
    Qs = [ssm.Q(L.T + t).numpy()[0] for t in range(Tf)]
    covs = [P]
    for t in range(Tf - 1):
        covs.append(Fs[t] @ covs[-1] @ Fs[t].T + Qs[t])
    C = np.zeros((Tf, Tf))
    for a in range(Tf):
        M = covs[a].copy()
        C[a, a] = H @ M @ H + float(ssm.obs_var[0])
        for b_ in range(a + 1, Tf):
            M = M @ Fs[b_ - 1].T
            C[a, b_] = C[b_, a] = H @ M @ H
    var_sum = C.sum()
    assert abs(s.sum(axis=1).var() / var_sum - 1.0) < 0.08
    emp = np.cov(s.T)
    assert np.abs(emp - C).max() < 0.12 * np.abs(C).max()


BIG = [(2, 3, 50, 9, 2, 8, 1, 0.0), (2, 2, 64, 11, 0, 9, 2, 0.1), (1, 3, 120, 60, 1, 52, 1, 0.0)]


@pytest.mark.parametrize('N,G,T,Tf,k,S,r,nan_frac', BIG)
def test_big_latent_path_filter_and_forecast(N, G, T, Tf, k, S, r, nan_frac):
    """nseasons outside the register-resident set: warp-per-lane kernels (ci_bigd.cu)."""
    L, prob, db = _setup(N, T, Tf, k, S, r, 31, nan_frac)
    B = N * G
    u = _draws(L, B, 7)
    ud = torch.as_tensor(u).cuda()
    pm, pv, fs = db.filter_predict(ud)
    series = torch.arange(B) // G
    mean_o, var_o, _, _, _, _ = O.one_step_predictive(prob, torch.as_tensor(u.T.astype(np.float64)), series)
    np.testing.assert_allclose(pm.cpu().numpy().T, mean_o.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pv.cpu().numpy().T, var_o.numpy(), rtol=1e-4)
    nsamp = 4000
    sims, mean = db.forecast_sample(ud, fs, nsamp, seed=5)
    means_o, vars_o = O.forecast_moments(prob, torch.as_tensor(u.T.astype(np.float64)), series)
    means_o = means_o.numpy().reshape(N, G, Tf)
    vars_o = vars_o.numpy().reshape(N, G, Tf)
    np.testing.assert_allclose(mean.cpu().numpy(), means_o.mean(axis=1), rtol=1e-4, atol=1e-4)
    s = sims.cpu().numpy().astype(np.float64)
    mix_mean = means_o.mean(axis=1)
    mix_var = (vars_o + means_o ** 2).mean(axis=1) - mix_mean ** 2
    assert np.all(np.abs(s.mean(axis=1) - mix_mean) < 5 * np.sqrt(mix_var / nsamp))
    assert np.all(np.abs(s.var(axis=1) / mix_var - 1.0) < 0.25)


GEN = [(O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 2, 3, 50, 9, 2, 1, 1, 0.0),
       (O.FLAG_LLT | O.FLAG_OBS_LOGNORMAL, 2, 2, 63, 10, 0, 7, 1, 0.1),
       (O.FLAG_DENSE_REG | O.FLAG_LEVEL_LOGNORMAL, 1, 3, 48, 8, 3, 4, 2, 0.0)]


@pytest.mark.parametrize('flags,N,G,T,Tf,k,S,r,nan_frac', GEN)
def test_wider_model_set_filter_and_forecast(flags, N, G, T, Tf, k, S, r, nan_frac):
    from tfcausalimpact_b200.engine import DeviceBatch
    L = O.Layout(T, Tf, k, S, r, flags)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=33, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r, flags=flags)
    B = N * G
    rng = np.random.default_rng(8)
    u = (rng.normal(size=(L.P, B)) * 0.4).astype(np.float32)
    u[0] -= 1.0
    u[1:1 + L.nt] -= 2.0
    ud = torch.as_tensor(u).cuda()
    pm, pv, fs = db.filter_predict(ud)
    series = torch.arange(B) // G
    uo = torch.as_tensor(u.T.astype(np.float64))
    mean_o, var_o, _, _, _, _ = O.one_step_predictive(prob, uo, series)
    np.testing.assert_allclose(pm.cpu().numpy().T, mean_o.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pv.cpu().numpy().T, var_o.numpy(), rtol=1e-4)
    nsamp = 4000
    sims, mean = db.forecast_sample(ud, fs, nsamp, seed=5)
    means_o, vars_o = O.forecast_moments(prob, uo, series)
    means_o = means_o.numpy().reshape(N, G, Tf)
    vars_o = vars_o.numpy().reshape(N, G, Tf)
    np.testing.assert_allclose(mean.cpu().numpy(), means_o.mean(axis=1), rtol=1e-4, atol=1e-4)
    s = sims.cpu().numpy().astype(np.float64)
    mix_mean = means_o.mean(axis=1)
    mix_var = (vars_o + means_o ** 2).mean(axis=1) - mix_mean ** 2
    assert np.all(np.abs(s.mean(axis=1) - mix_mean) < 5 * np.sqrt(mix_var / nsamp))
    assert np.all(np.abs(s.var(axis=1) / mix_var - 1.0) < 0.25)


TWO_SEASONAL = [(O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 1, 3, 120, 12, 3, 4, 1, 52, 1, 0.0),
                (0, 2, 2, 60, 14, 2, 7, 1, 4, 3, 0.1),
                (O.FLAG_LLT | O.FLAG_OBS_LOGNORMAL, 2, 2, 50, 11, 0, 3, 2, 5, 1, 0.0)]


@pytest.mark.parametrize('flags,N,G,T,Tf,k,S,r,S2,r2,nan_frac', TWO_SEASONAL)
def test_two_seasonal_components_filter_and_forecast(flags, N, G, T, Tf, k, S, r, S2, r2, nan_frac):
    from tfcausalimpact_b200.engine import DeviceBatch
    L = O.Layout(T, Tf, k, S, r, flags, S2, r2)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=35, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r, flags=flags,
                     nseasons2=S2, season_duration2=r2)
    B = N * G
    rng = np.random.default_rng(8)
    u = (rng.normal(size=(L.P, B)) * 0.4).astype(np.float32)
    u[0] -= 1.0
    u[1:1 + L.nt] -= 2.0
    ud = torch.as_tensor(u).cuda()
    pm, pv, fs = db.filter_predict(ud)
    series = torch.arange(B) // G
    uo = torch.as_tensor(u.T.astype(np.float64))
    mean_o, var_o, _, _, _, _ = O.one_step_predictive(prob, uo, series)
    np.testing.assert_allclose(pm.cpu().numpy().T, mean_o.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pv.cpu().numpy().T, var_o.numpy(), rtol=1e-4)
    nsamp = 4000
    sims, mean = db.forecast_sample(ud, fs, nsamp, seed=5)
    means_o, vars_o = O.forecast_moments(prob, uo, series)
    means_o = means_o.numpy().reshape(N, G, Tf)
    vars_o = vars_o.numpy().reshape(N, G, Tf)
    np.testing.assert_allclose(mean.cpu().numpy(), means_o.mean(axis=1), rtol=1e-4, atol=1e-4)
    s = sims.cpu().numpy().astype(np.float64)
    mix_mean = means_o.mean(axis=1)
    mix_var = (vars_o + means_o ** 2).mean(axis=1) - mix_mean ** 2
    assert np.all(np.abs(s.mean(axis=1) - mix_mean) < 5 * np.sqrt(mix_var / nsamp))
    assert np.all(np.abs(s.var(axis=1) / mix_var - 1.0) < 0.25)
