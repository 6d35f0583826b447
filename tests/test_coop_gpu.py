"""GPU parity of the lane-cooperative evaluation kernel (csrc/ci_eval_coop.cu: 8 / 4 sub-threads per lane for
batches that cannot fill the GPU at one thread per lane) against the oracle (dense filter + autograd, fp64) and
against the thread-per-lane kernel.  Tolerance as for every evaluation kernel: log-prob rtol 1e-4; gradient per
component rtol 1e-4 with atol 1e-5 * max|g| (north_star / SURVEY 8(d))."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import O, synth_problem

pytestmark = pytest.mark.gpu


def _set(tpl, max_lanes=-1, mma=1):
    from tfcausalimpact_b200 import _native as nt
    lib = nt.lib()
    nt.check(lib.ci_set_option(b'coop_threads_per_lane', ctypes.c_long(tpl)), 'ci_set_option')
    nt.check(lib.ci_set_option(b'coop_max_lanes', ctypes.c_long(max_lanes)), 'ci_set_option')
    nt.check(lib.ci_set_option(b'coop_mma', ctypes.c_long(mma)), 'ci_set_option')


@pytest.fixture(autouse=True)
def _restore_default():
    yield
    _set(-1, -1, 1)


COOP_CASES = [
    # S, k, T, nan_frac, N, G
    (7, 10, 365, 0.0, 5, 8),     # cfg3 shape
    (7, 50, 365, 0.05, 3, 8),    # cfg4 shape with missing steps
    (7, 5, 61, 0.05, 7, 6),      # series straddling warps, ragged last chunk
    (7, 3, 33, 0.1, 9, 1),       # one lane per series: 4 series per warp at 8 sub-threads, padding lanes
    (7, 0, 40, 0.0, 3, 3),       # no regression
    (7, 1, 30, 0.0, 2, 5),       # fewer covariates than sub-threads
    (2, 4, 50, 0.1, 4, 3),
    (3, 9, 41, 0.0, 11, 5),
    (4, 2, 50, 0.1, 4, 3),
    (5, 17, 37, 0.0, 6, 2),
    (6, 8, 45, 0.3, 5, 4),
    (7, 50, 5, 0.0, 3, 8),       # T < a season cycle / the tape prefetch distance
    (7, 4, 9, 0.0, 1, 1),        # a single lane
    # 8 / 16 chains per series: at 4 sub-threads the regression passes run on the tensor cores (3xTF32)
    (7, 64, 33, 0.1, 2, 8),      # widest regression, odd number of 4-step chunks (ragged last 8-step group)
    (7, 1, 30, 0.0, 2, 16),      # one covariate, two warps per series
    (5, 3, 41, 0.0, 3, 8),
    (2, 9, 9, 0.0, 1, 8),
]


@pytest.mark.parametrize('tpl', [8, 4])
@pytest.mark.parametrize('S,k,T,nan_frac,N,G', COOP_CASES)
def test_coop_logp_grad_matches_oracle(tpl, S, k, T, nan_frac, N, G):
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 6
    L = O.Layout(T, Tf, k, S, 1)
    y, _, X = synth_problem(N, T, Tf, k, S, 1, seed=S * 100 + k + T, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S)
    rng = np.random.default_rng(tpl)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.7).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    ud = torch.as_tensor(u).cuda()
    _set(tpl)
    lp, g = db.logp_grad(ud)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4, atol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 2e-5 * gmax)
    # and against the thread-per-lane kernel on the same inputs (same formulation, different summation order)
    _set(0)
    lp1, g1 = db.logp_grad(ud)
    np.testing.assert_allclose(lp, lp1.cpu().numpy(), rtol=2e-5, atol=2e-4)
    assert np.all(np.abs(g - g1.cpu().numpy()) <= 2e-4 * np.abs(go) + 4e-5 * gmax)


@pytest.mark.parametrize('S,k,T,nan_frac,N,G', [c for c in COOP_CASES if c[5] % 8 == 0])
def test_coop_tensor_core_passes_match_ffma_passes(S, k, T, nan_frac, N, G):
    """The 3xTF32 tensor-core regression passes against the FFMA2 passes of the same kernel: same inputs, the two
    differ only in the summation order / the dropped lo*lo term of the split (2^-22 relative)."""
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 6
    L = O.Layout(T, Tf, k, S, 1)
    y, _, X = synth_problem(N, T, Tf, k, S, 1, seed=S * 100 + k + T, nan_frac=nan_frac)
    db = DeviceBatch(y, X if k > 0 else None, Tf=Tf, nseasons=S)
    rng = np.random.default_rng(3)
    u = (rng.normal(size=(L.P, N * G)) * 0.7).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    ud = torch.as_tensor(u).cuda()
    out = {}
    for mma in (1, 0):
        _set(4, -1, mma)
        lp, g = db.logp_grad(ud)
        out[mma] = (lp.cpu().numpy().astype(np.float64), g.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(out[1][0], out[0][0], rtol=2e-5, atol=2e-4)
    gmax = np.abs(out[0][1]).max(axis=0, keepdims=True)
    assert np.all(np.abs(out[1][1] - out[0][1]) <= 5e-5 * np.abs(out[0][1]) + 2e-5 * gmax)
    assert np.abs(out[1][1] - out[0][1]).max() > 0.0      # the two paths really are different code


@pytest.mark.parametrize('tpl', [8, 4])
def test_coop_hmc_trajectory_matches_oracle(tpl):
    """The fused leapfrog kick / drift of the cooperative kernel: shared-Philox HMC trajectories vs the oracle."""
    from tfcausalimpact_b200.engine import DeviceBatch
    N, G, T, k, S = 3, 3, 45, 5, 7
    L = O.Layout(T, 5, k, S, 1)
    y, _, X = synth_problem(N, T, 5, k, S, 1, seed=21)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    db = DeviceBatch(y, X, Tf=5, nseasons=S)
    B = N * G
    rng = np.random.default_rng(5)
    u0 = (rng.normal(size=(L.P, B)) * 0.3).astype(np.float32)
    u0[0] -= 1.0
    u0[1] -= 2.0
    step0 = np.full((L.P, B), 0.02, dtype=np.float32) * (1 + rng.random(size=(L.P, B))).astype(np.float32)
    nres, nwarm, nadapt, leap, seed = 4, 6, 4, 5, 77
    _set(tpl)
    u_dev = torch.as_tensor(u0).cuda()
    draws, stats = db.hmc_run(u_dev, torch.as_tensor(step0).cuda(), nres, nwarm, nadapt, leap, 0.75, seed)
    dr_o, acc_o, fac_o = O.hmc_run(prob, np.arange(B) // G, torch.as_tensor(u0.T.astype(np.float64)),
                                   torch.as_tensor(step0.T.astype(np.float64)), nres, nwarm, leap, seed, 0.75, nadapt)
    dr = draws.cpu().numpy()
    np.testing.assert_allclose(np.transpose(dr, (0, 2, 1)), dr_o.numpy(), rtol=2e-3, atol=4e-4)   # 10 transitions amplify fp32 rounding differences
    np.testing.assert_allclose(stats.cpu().numpy()[2], fac_o.numpy(), rtol=2e-3)


def test_coop_multi_wave_and_large_batch_agree_with_thread_per_lane():
    """Size-independent check at a sub-wave batch of the named workload (1,250 series x 8 chains = the 8-GPU
    split of configs[3]): both kernels give the same log-prob and gradient."""
    from tfcausalimpact_b200.engine import DeviceBatch
    N, G, T, Tf, k, S = 1250, 8, 365, 30, 50, 7
    y, _, X = synth_problem(N, T, Tf, k, S, 1, seed=5)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    g = torch.Generator().manual_seed(0)
    u = (torch.randn(db.P, N * G, generator=g) * 0.3).cuda()
    u[0] -= 1.0
    u[1] -= 2.0
    _set(0)
    lp0, g0 = db.logp_grad(u)
    lp0, g0 = lp0.clone(), g0.clone()
    for tpl in (8, 4):
        _set(tpl, 1 << 30)
        lp, gr = db.logp_grad(u)
        assert torch.isfinite(lp).all() and torch.isfinite(gr).all()
        torch.testing.assert_close(lp, lp0, rtol=2e-5, atol=1e-3)
        gmax = g0.abs().amax(dim=0, keepdim=True)
        assert bool(((gr - g0).abs() <= 2e-4 * g0.abs() + 4e-5 * gmax).all())
