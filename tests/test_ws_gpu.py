"""GPU parity of the warp-specialised thread-per-lane evaluation kernel (csrc/ci_eval_ws.cu: the filter / adjoint steps
in one warp, the regression of the same 32 lanes one chunk ahead in a second warp, for mid-size batches) against the
oracle (dense filter + autograd, fp64) and against the single-warp kernel it replaces there.  Tolerance as for every
evaluation kernel: log-prob rtol 1e-4; gradient per component rtol 1e-4 with atol 1e-5 * max|g|."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import O, synth_problem

pytestmark = pytest.mark.gpu


def _set(ws, coop=-1):
    from tfcausalimpact_b200 import _native as nt
    lib = nt.lib()
    nt.check(lib.ci_set_option(b'eval_ws', ctypes.c_long(ws)), 'ci_set_option')
    nt.check(lib.ci_set_option(b'coop_threads_per_lane', ctypes.c_long(coop)), 'ci_set_option')


@pytest.fixture(autouse=True)
def _restore_default():
    yield
    _set(1, -1)


WS_CASES = [
    # S, k, T, nan_frac, N, G
    (7, 10, 365, 0.0, 5, 8),     # cfg3 shape
    (7, 50, 365, 0.05, 3, 8),    # cfg4 shape with missing steps
    (7, 5, 61, 0.05, 7, 6),      # series straddling warps, ragged last chunk
    (3, 9, 41, 0.0, 11, 5),      # partial last warp (padding lanes shadow the last lane)
    (5, 17, 37, 0.0, 6, 4),
    (2, 4, 50, 0.1, 4, 4),       # smallest regression the kernel takes
    (7, 50, 5, 0.0, 3, 8),       # two chunks only
    (7, 4, 3, 0.0, 1, 4),        # a single chunk
    (6, 8, 45, 0.3, 5, 4),
    (4, 6, 30, 0.0, 40, 8),      # ten warps
    (7, 64, 33, 0.1, 9, 32),     # one series per warp
    (7, 100, 40, 0.0, 3, 8),     # more covariates than the cooperative kernel takes: this kernel serves small batches too
]


@pytest.mark.parametrize('S,k,T,nan_frac,N,G', WS_CASES)
def test_ws_logp_grad_matches_oracle(S, k, T, nan_frac, N, G):
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 6
    L = O.Layout(T, Tf, k, S, 1)
    y, _, X = synth_problem(N, T, Tf, k, S, 1, seed=S * 100 + k + T, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    rng = np.random.default_rng(5)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.7).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    ud = torch.as_tensor(u).cuda()
    _set(1, 0)                                   # cooperative kernel off: the warp-specialised kernel serves the call
    lp, g = db.logp_grad(ud)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4, atol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 2e-5 * gmax)
    # the single-warp kernel runs the same per-lane arithmetic in the same order: results agree to rounding of the
    # regression sums (four-way instead of five-way batches)
    _set(0, 0)
    lp1, g1 = db.logp_grad(ud)
    np.testing.assert_allclose(lp, lp1.cpu().numpy(), rtol=2e-5, atol=2e-4)
    assert np.all(np.abs(g - g1.cpu().numpy()) <= 2e-4 * np.abs(go) + 4e-5 * gmax)


def test_ws_hmc_trajectory_matches_single_warp_kernel():
    """The fused leapfrog kick / drift behind the warp-specialised kernel: identical Philox streams, so whole HMC
    trajectories must agree with the single-warp kernel's to rounding."""
    from tfcausalimpact_b200.engine import DeviceBatch
    N, G, T, Tf, k, S = 6, 8, 90, 6, 12, 7
    y, _, X = synth_problem(N, T, Tf, k, S, 1, seed=77)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    g = torch.Generator(device='cpu').manual_seed(1)
    u0 = (torch.randn(db.P, N * G, generator=g) * 0.2).cuda()
    step0 = torch.full_like(u0, 0.01)
    out = {}
    for ws in (1, 0):
        _set(ws, 0)
        draws, stats = db.hmc_run(u0.clone(), step0, 6, 0, 0, 10, 0.75, seed=4)
        out[ws] = (draws.cpu().numpy(), float(stats[0].mean()))
    np.testing.assert_allclose(out[1][0], out[0][0], rtol=2e-3, atol=2e-3)
    assert abs(out[1][1] - out[0][1]) < 0.05
