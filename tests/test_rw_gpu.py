"""GPU parity of the register-resident few-lane / large-latent evaluation kernel (csrc/ci_eval_rw.cu: 4 or 8 warps
per lane, covariance and adjoint in registers) against the oracle (dense filter + autograd, fp64) and against the
shared-memory CTA-per-lane kernel it replaces (k_eval_cta, ci_bigd.cu).  Tolerance as for every evaluation kernel:
log-prob rtol 1e-4; gradient per component rtol 1e-4 with atol 1e-5 * max|g| (north_star / SURVEY 8(d))."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import O, synth_problem

pytestmark = pytest.mark.gpu


def _set(rw, warps=8):
    from tfcausalimpact_b200 import _native as nt
    lib = nt.lib()
    nt.check(lib.ci_set_option(b'eval_rw', ctypes.c_long(rw)), 'ci_set_option')
    nt.check(lib.ci_set_option(b'eval_rw_warps', ctypes.c_long(warps)), 'ci_set_option')


@pytest.fixture(autouse=True)
def _restore_default():
    yield
    _set(1, 4)


RW_CASES = [
    # S, r, k, T, nan_frac, N, G
    (52, 1, 3, 130, 0.0, 2, 2),    # configs[4] latent size
    (52, 1, 0, 200, 0.1, 3, 1),    # no regression, masked steps, one lane per series
    (52, 7, 2, 400, 0.0, 1, 1),    # weekly blocks of a yearly season, a single lane
    (63, 1, 5, 90, 0.05, 2, 1),    # largest latent the kernel takes (D = 64)
    (8, 1, 2, 60, 0.0, 3, 2),      # smallest (D = 9)
    (9, 2, 0, 75, 0.1, 2, 3),
    (10, 3, 40, 90, 0.05, 2, 2),
    (24, 1, 70, 100, 0.0, 2, 2),   # more covariates than threads per time slice allow (k > NT / 2)
    (31, 1, 1, 33, 0.0, 5, 4),     # D = 32: exactly one column per lane
    (32, 2, 1, 70, 0.2, 4, 2),     # D = 33: first latent with a second column slot
    (40, 1, 4, 7, 0.0, 2, 2),      # T shorter than one season cycle
]


@pytest.mark.parametrize('warps', [4, 8])
@pytest.mark.parametrize('S,r,k,T,nan_frac,N,G', RW_CASES)
def test_rw_logp_grad_matches_oracle(warps, S, r, k, T, nan_frac, N, G):
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 6
    L = O.Layout(T, Tf, k, S, r)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=S * 100 + k + T, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S, season_duration=r)
    rng = np.random.default_rng(warps)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.7).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    ud = torch.as_tensor(u).cuda()
    _set(1, warps)
    lp, g = db.logp_grad(ud)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4, atol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 2e-5 * gmax)
    # and against the shared-memory CTA-per-lane kernel on the same inputs
    _set(0)
    lp1, g1 = db.logp_grad(ud)
    np.testing.assert_allclose(lp, lp1.cpu().numpy(), rtol=2e-5, atol=2e-4)
    assert np.all(np.abs(g - g1.cpu().numpy()) <= 2e-4 * np.abs(go) + 4e-5 * gmax)
