"""GPU parity: C-ABI ci_kalman_logp_grad vs the oracle (dense filter + autograd, fp64).
Tolerance (north_star): rtol 1e-4 on log-prob; gradient per component rtol 1e-4 with
atol 1e-5 * max|g|."""
import numpy as np
import pytest
import torch

from helpers import O, sconst_from_oracle, synth_problem

pytestmark = pytest.mark.gpu

CASES = [
    (1, 1, 0, 40, 0.0, 1, 1),
    (1, 1, 1, 70, 0.0, 1, 1),      # arma-like (cfg1)
    (1, 1, 3, 90, 0.0, 1, 1),      # comparison-like (cfg2)
    (7, 1, 10, 365, 0.0, 5, 8),    # cfg3 shape
    (7, 1, 50, 365, 0.05, 3, 8),   # cfg4 shape, with missing steps
    (4, 3, 2, 50, 0.1, 4, 3),
    (2, 1, 1, 33, 0.0, 2, 2),
    (12, 2, 4, 100, 0.0, 2, 2),
    (3, 1, 0, 20, 0.0, 7, 5),
    (7, 1, 5, 61, 0.05, 7, 6),     # staged path, series straddling warps (G = 6), ragged last chunk
    (3, 1, 9, 41, 0.0, 11, 5),     # staged path, G = 5, partial last warp
    (5, 2, 6, 50, 0.1, 4, 12),     # staged residual-basis path (season_duration 2)
    # nseasons outside the register-resident set -> warp-per-lane shared-memory path (ci_bigd.cu)
    (8, 1, 2, 60, 0.0, 3, 2),
    (9, 2, 0, 75, 0.1, 2, 3),
    (10, 3, 40, 90, 0.05, 2, 2),
    (52, 1, 3, 130, 0.0, 2, 2),   # configs[4] latent size (few lanes -> CTA-per-lane kernel)
    # fewer than 4 lanes per series with >= 8 series: covariate-major design block (ci_data.d_XT)
    (7, 1, 6, 50, 0.05, 9, 1),
    (5, 1, 50, 37, 0.0, 12, 2),
    (7, 1, 3, 61, 0.1, 33, 3),
    (1, 1, 4, 30, 0.0, 40, 1),
    (6, 3, 5, 45, 0.1, 10, 2),     # residual-basis path (season_duration 3)
    # the same large-latent layouts with enough lanes (B > 296) to stay on the warp-per-lane kernel
    (9, 2, 2, 40, 0.1, 40, 8),
    (52, 1, 0, 60, 0.0, 50, 6),
]


@pytest.mark.parametrize('S,r,k,T,nan_frac,N,G', CASES)
def test_logp_grad_matches_oracle(S, r, k, T, nan_frac, N, G):
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 9
    L = O.Layout(T, Tf, k, S, r)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=S * 100 + k, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r)
    # device statistics vs oracle statistics
    np.testing.assert_allclose(db.sconst.cpu().numpy(), sconst_from_oracle(prob), rtol=2e-5, atol=2e-6)
    rng = np.random.default_rng(1)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.7).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    lp, g = db.logp_grad(torch.as_tensor(u).cuda())
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 1e-5 * gmax)


@pytest.mark.parametrize('N', [128, 512, 1250, 2500, 7200])
def test_linearity_property_full_size(N):
    """Size-independent property at cfg4 scale: with all variance parameters fixed, dL/dr is an
    affine function of the residual series, so grad wrt weights_noncentered is affine along a
    line in weight space (checked at three collinear points).  The series counts put the call on each evaluation
    kernel of the default dispatch (8 chains): cooperative with 8 and with 4 sub-threads (1,024 / 4,096 / 10,000 lanes:
    the 8-GPU share of configs[3]), warp-specialised (20,000 lanes: the 4-GPU share), thread-per-lane (57,600 lanes)."""
    from tfcausalimpact_b200.engine import DeviceBatch
    G, T, Tf, k, S = 8, 365, 30, 50, 7
    rng = np.random.default_rng(3)
    y = rng.normal(size=(N, T)).astype(np.float32).cumsum(axis=1) * 0.1 + rng.normal(size=(N, T)).astype(np.float32)
    X = rng.normal(size=(N, T + Tf, k)).astype(np.float32)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    L = O.Layout(T, Tf, k, S, 1)
    g = torch.Generator(device='cpu').manual_seed(0)
    u0 = (torch.randn(L.P, N * G, generator=g) * 0.3)
    d = torch.zeros_like(u0)
    d[L.i_wn:L.i_wn + k] = torch.randn(k, N * G, generator=g) * 0.2
    grads = []
    for a in (0.0, 1.0, 2.0):
        _, gr = db.logp_grad((u0 + a * d).cuda())
        grads.append(gr[L.i_wn:L.i_wn + k].double().cpu())
    mid = 0.5 * (grads[0] + grads[2])
    scale = grads[1].abs().max()
    assert torch.all(torch.isfinite(grads[1]))
    assert (mid - grads[1]).abs().max() <= 2e-4 * scale


# wider model set (custom `model=`): LocalLinearTrend, dense LinearRegression, tfp.sts default priors
GEN_CASES = [
    # flags, S, r, k, T, nan_frac, N, G
    (O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 1, 1, 3, 60, 0.0, 3, 2),      # notebook cell 44 shape
    (O.FLAG_LLT | O.FLAG_OBS_LOGNORMAL, 7, 1, 0, 70, 0.1, 2, 3),                         # main.py:191-193 shape
    (O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL | O.FLAG_LEVEL_LOGNORMAL, 4, 2, 4, 50, 0.05, 2, 2),
    (O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 12, 1, 2, 64, 0.0, 2, 2),
    (O.FLAG_OBS_LOGNORMAL, 1, 1, 2, 45, 0.0, 3, 1),                                      # sparse + LogNormal noise prior
    (O.FLAG_LLT, 3, 1, 40, 40, 0.0, 1, 2),
    # generic priors WITHOUT a trend slope and a large latent, few lanes: the register-resident 4-warp kernel (ci_eval_rw.cu)
    (O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL | O.FLAG_LEVEL_LOGNORMAL, 12, 1, 3, 64, 0.05, 2, 2),
    (O.FLAG_OBS_LOGNORMAL, 20, 2, 0, 50, 0.0, 2, 1),
    (O.FLAG_DENSE_REG, 52, 1, 2, 110, 0.0, 1, 2),
]


@pytest.mark.parametrize('flags,S,r,k,T,nan_frac,N,G', GEN_CASES)
def test_wider_model_set_logp_grad_matches_oracle(flags, S, r, k, T, nan_frac, N, G):
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 7
    L = O.Layout(T, Tf, k, S, r, flags)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=S * 10 + k + flags, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r, flags=flags)
    assert db.P == L.P
    rng = np.random.default_rng(2)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.6).astype(np.float32)
    u[0] -= 1.0
    u[1:1 + L.nt] -= 2.0
    lp, g = db.logp_grad(torch.as_tensor(u).cuda())
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 1e-5 * gmax)
    # constrain / unconstrain round trip of the generic parameter block
    ud = torch.as_tensor(u).cuda()
    th = db.constrain(ud)
    np.testing.assert_allclose(db.unconstrain(th).cpu().numpy(), u, rtol=2e-4, atol=2e-4)


# two Seasonal components (custom `model=`; notebooks/getting_started.ipynb cell 116: Month(4) + Year(52))
TWO_SEASONAL_CASES = [
    # flags, S, r, S2, r2, k, T, nan_frac, N, G
    (O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 4, 1, 52, 1, 3, 120, 0.0, 2, 2),   # the notebook's model
    (0, 7, 1, 4, 3, 2, 60, 0.1, 2, 2),                   # default priors + a second seasonal with duration 3
    (O.FLAG_LLT | O.FLAG_OBS_LOGNORMAL, 3, 2, 5, 1, 0, 50, 0.05, 1, 3),
    (O.FLAG_OBS_LOGNORMAL | O.FLAG_LEVEL_LOGNORMAL, 12, 1, 30, 2, 1, 90, 0.0, 1, 2),
    (O.FLAG_LLT | O.FLAG_DENSE_REG | O.FLAG_OBS_LOGNORMAL, 4, 1, 9, 1, 2, 36, 0.05, 50, 6),   # B = 300: warp-per-lane kernel
]


@pytest.mark.parametrize('flags,S,r,S2,r2,k,T,nan_frac,N,G', TWO_SEASONAL_CASES)
def test_two_seasonal_components_logp_grad_matches_oracle(flags, S, r, S2, r2, k, T, nan_frac, N, G):
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 7
    L = O.Layout(T, Tf, k, S, r, flags, S2, r2)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=S * 10 + S2 + k + flags, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r, flags=flags,
                     nseasons2=S2, season_duration2=r2)
    assert db.P == L.P
    rng = np.random.default_rng(2)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.6).astype(np.float32)
    u[0] -= 1.0
    u[1:1 + L.nt] -= 2.0
    lp, g = db.logp_grad(torch.as_tensor(u).cuda())
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 1e-5 * gmax)
    ud = torch.as_tensor(u).cuda()
    th = db.constrain(ud)
    np.testing.assert_allclose(db.unconstrain(th).cpu().numpy(), u, rtol=2e-4, atol=2e-4)


# edge shapes: series shorter than a staging chunk / the tape prefetch distance / one season cycle, single
# lane, single covariate, a masked run at the very start after the first observation
EDGE_CASES = [
    # S, r, k, T, N, G, masked steps
    (1, 1, 0, 2, 1, 1, ()),
    (1, 1, 1, 3, 2, 1, (1,)),
    (7, 1, 50, 5, 3, 8, (1, 2)),          # staged effects path, T < 8 (no tape prefetch), T < S
    (7, 1, 4, 9, 1, 4, (3,)),
    (2, 1, 2, 4, 2, 2, ()),
    (3, 2, 1, 7, 1, 3, (1, 2, 3)),
    (52, 1, 0, 6, 1, 1, ()),              # CTA-per-lane kernel, T << S
    (9, 1, 3, 5, 40, 8, (2,)),            # warp-per-lane kernel
]


@pytest.mark.parametrize('S,r,k,T,N,G,masked', EDGE_CASES)
def test_edge_shapes_match_oracle(S, r, k, T, N, G, masked):
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 3
    L = O.Layout(T, Tf, k, S, r)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=900 + S + T)
    for t in masked:
        y[:, t] = np.nan
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r)
    rng = np.random.default_rng(4)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.5).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    lp, g = db.logp_grad(torch.as_tensor(u).cuda())
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4, atol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 2e-5 * gmax)
    # the predictive path on the same tiny shapes
    ud = torch.as_tensor(u).cuda()
    pm, pv, fs = db.filter_predict(ud)
    mean_o, var_o, _, _, _, _ = O.one_step_predictive(prob, torch.as_tensor(u.T.astype(np.float64)), series)
    np.testing.assert_allclose(pm.cpu().numpy().T, mean_o.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pv.cpu().numpy().T, var_o.numpy(), rtol=1e-4)
    sims, mean = db.forecast_sample(ud, fs, 64, seed=1)
    assert torch.isfinite(sims).all() and torch.isfinite(mean).all()


def _fuzz_cases(n=24, seed=2024):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        S = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 7, 7, 12, 9]))
        r = int(rng.choice([1, 1, 1, 2, 3])) if S > 1 else 1
        k = int(rng.choice([0, 1, 3, 4, 5, 9, 10, 17, 50]))
        T = int(rng.integers(6, 70))
        G = int(rng.choice([1, 2, 3, 4, 5, 8, 8, 12]))
        N = int(rng.choice([1, 2, 5, 8, 9, 13, 33]))
        nan_frac = float(rng.choice([0.0, 0.0, 0.1, 0.3]))
        out.append((S, r, k, T, nan_frac, N, G))
    return out


@pytest.mark.parametrize('S,r,k,T,nan_frac,N,G', _fuzz_cases())
def test_random_shapes_match_oracle(S, r, k, T, nan_frac, N, G):
    """Seeded random layouts across every evaluation kernel and staging policy (direct / TMA per warp / cp.async
    per lane, effects and residual basis, capped and uncapped registers, warp- and CTA-per-lane)."""
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 5
    L = O.Layout(T, Tf, k, S, r)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=S * 1000 + k * 10 + T, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r)
    rng = np.random.default_rng(T)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.6).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    lp, g = db.logp_grad(torch.as_tensor(u).cuda())
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4, atol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 2e-5 * gmax)
