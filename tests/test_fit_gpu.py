"""GPU parity of the fit drivers (C-ABI ci_vi_fit / ci_vi_draw / ci_hmc_run) against the oracle's
restatement of tfp.vi.fit_surrogate_posterior / tfp.sts.fit_with_hmc (oracle/sts_oracle.py).
Device and oracle share the Philox streams, so trajectories agree up to fp32 rounding (the oracle
runs in fp64); tolerances are stated per test."""
import numpy as np
import pytest
import torch

from helpers import O, synth_problem

pytestmark = pytest.mark.gpu


def _setup(N, T, Tf, k, S, r=1, seed=0, nan_frac=0.0):
    from tfcausalimpact_b200.engine import DeviceBatch
    L = O.Layout(T, Tf, k, S, r)
    y, ypost, X = synth_problem(N, T, Tf, k, S, r, seed=seed, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r)
    return L, prob, db


@pytest.mark.parametrize('N,T,k,S,units,ssize', [(3, 40, 2, 4, 1, 1), (2, 30, 0, 1, 2, 3), (2, 50, 3, 7, 1, 5)])
def test_vi_matches_oracle(N, T, k, S, units, ssize):
    L, prob, db = _setup(N, T, 5, k, S, seed=7)
    steps, seed = 25, 1234
    mu, rho, loss = db.vi_fit(units, ssize, steps, 0.1, 0.01, seed, return_loss=True)
    series_of_unit = np.repeat(np.arange(N), units)
    mu_o, rho_o, loss_o = O.vi_fit(prob, series_of_unit, steps, ssize, seed)
    # first-step loss is a pure function of the shared random numbers: tight
    np.testing.assert_allclose(loss.cpu().numpy()[0], loss_o[0], rtol=2e-4)
    # Adam's normalised steps keep fp32/fp64 trajectories together over 25 steps
    np.testing.assert_allclose(mu.cpu().numpy().T, mu_o.numpy(), atol=5e-3, rtol=5e-3)
    np.testing.assert_allclose(rho.cpu().numpy().T, rho_o.numpy(), atol=5e-3, rtol=5e-3)
    # surrogate.sample(): same Philox stream
    dr = db.vi_draw(mu, rho, 7, seed=99).cpu().numpy()           # [P, U*7]
    dr_o = O.vi_draw(torch.as_tensor(mu.cpu().numpy().T.astype(np.float64)),
                     torch.as_tensor(rho.cpu().numpy().T.astype(np.float64)), 7, 99).numpy()  # [U, 7, P]
    np.testing.assert_allclose(dr.T.reshape(dr_o.shape), dr_o, rtol=1e-4, atol=1e-5)


def test_vi_reduces_loss_and_recovers_weights():
    """200 Adam steps (the reference's setting, model.py:369) drive the negative ELBO down."""
    L, prob, db = _setup(4, 120, 10, 3, 1, seed=3)
    mu, rho, loss = db.vi_fit(1, 1, 200, 0.1, 0.01, 5, return_loss=True)
    loss = loss.cpu().numpy()
    assert np.all(np.isfinite(loss))
    assert np.all(loss[-20:].mean(axis=0) < loss[:5].mean(axis=0) - 10.0)


@pytest.mark.parametrize('N,G,T,k,S,r', [(3, 2, 40, 2, 4, 1), (2, 3, 30, 0, 1, 1), (2, 2, 45, 1, 3, 2)])
def test_hmc_matches_oracle(N, G, T, k, S, r):
    L, prob, db = _setup(N, T, 5, k, S, r, seed=11)
    B = N * G
    rng = np.random.default_rng(5)
    u0 = (rng.normal(size=(L.P, B)) * 0.3).astype(np.float32)
    u0[0] -= 1.0
    u0[1] -= 2.0
    step0 = np.full((L.P, B), 0.02, dtype=np.float32) * (1 + rng.random(size=(L.P, B))).astype(np.float32)
    nres, nwarm, nadapt, leap, seed = 4, 6, 4, 5, 77
    u_dev = torch.as_tensor(u0).cuda()
    draws, stats = db.hmc_run(u_dev, torch.as_tensor(step0).cuda(), nres, nwarm, nadapt, leap, 0.75, seed)
    series_of_chain = np.arange(B) // G
    dr_o, acc_o, fac_o = O.hmc_run(prob, series_of_chain, torch.as_tensor(u0.T.astype(np.float64)),
                                   torch.as_tensor(step0.T.astype(np.float64)), nres, nwarm, leap, seed,
                                   0.75, nadapt)
    dr = draws.cpu().numpy()                              # [nres, P, B]
    np.testing.assert_allclose(np.transpose(dr, (0, 2, 1)), dr_o.numpy(), rtol=2e-3, atol=2e-4)
    st = stats.cpu().numpy()
    np.testing.assert_allclose(st[0], acc_o[nwarm:].mean(axis=0), atol=1e-6)
    np.testing.assert_allclose(st[1], acc_o[:nwarm].mean(axis=0), atol=1e-6)
    np.testing.assert_allclose(st[2], fac_o.numpy(), rtol=2e-3)
    # final state == last draw
    np.testing.assert_allclose(u_dev.cpu().numpy(), dr[-1], rtol=0, atol=0)


def test_hmc_energy_conservation_small_steps():
    """Size-independent property: with tiny steps the leapfrog integrator conserves H, so every
    proposal is accepted and the state moves."""
    L, prob, db = _setup(64, 100, 5, 5, 7, seed=2)
    B = 64 * 4
    g = torch.Generator().manual_seed(0)
    u = (torch.randn(L.P, B, generator=g) * 0.2).cuda()
    u0 = u.clone()
    step0 = torch.full((L.P, B), 1e-3).cuda()
    draws, stats = db.hmc_run(u, step0, 5, 0, 0, 10, 0.75, 3)
    assert float(stats[0].min()) == 1.0
    assert float((draws[-1] - u0).abs().max()) > 1e-3
