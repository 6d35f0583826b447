"""CPU check of the kernels' per-lane arithmetic (structured filter + hand-derived adjoint +
parameter VJP) against the oracle's dense TFP-style filter with autograd, through the TEST-ONLY
host shim.  Tolerance: north_star rtol 1e-4 (fp32 kernel arithmetic vs fp64 oracle)."""
import numpy as np
import pytest
import torch

from helpers import O, pack_X, sconst_from_oracle, shim_logp_grad, synth_problem

CASES = [
    # S, r, k, T, nan_frac
    (1, 1, 0, 40, 0.0),
    (1, 1, 3, 70, 0.1),
    (7, 1, 10, 120, 0.0),
    (7, 1, 50, 365, 0.05),
    (4, 3, 2, 50, 0.1),
    (2, 1, 1, 33, 0.0),
    (12, 2, 4, 100, 0.0),
    (3, 1, 0, 20, 0.0),
    (5, 1, 1, 64, 0.3),
    (6, 4, 0, 90, 0.0),
]


@pytest.mark.parametrize('S,r,k,T,nan_frac', CASES)
def test_lane_logp_grad_matches_oracle(S, r, k, T, nan_frac):
    N, G, Tf = 3, 4, 7
    L = O.Layout(T, Tf, k, S, r)
    y, _, X = synth_problem(N, T, Tf, k, S, r, seed=S * 100 + k, nan_frac=nan_frac)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    sc = sconst_from_oracle(prob)
    XY, Tpad = pack_X(X, y)
    rng = np.random.default_rng(1)
    B = N * G
    u = (rng.normal(size=(L.P, B)) * 0.7).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    lp, g = shim_logp_grad(L, XY, Tpad, sc, u, G)
    series = torch.arange(B) // G
    lpo, go = O.target_log_prob_and_grad(prob, torch.tensor(u.T.astype(np.float64)), series)
    lpo, go = lpo.numpy(), go.numpy().T
    np.testing.assert_allclose(lp, lpo, rtol=1e-4)
    gmax = np.abs(go).max(axis=0, keepdims=True)
    assert np.all(np.abs(g - go) <= 1e-4 * np.abs(go) + 1e-5 * gmax)


def test_oracle_gradient_vs_finite_differences():
    L = O.Layout(30, 5, 2, 4, 2)
    y, _, X = synth_problem(2, 30, 5, 2, 4, 2, seed=5, nan_frac=0.1)
    prob = O.Problem.build(y, X, L, dtype=torch.float64)
    rng = np.random.default_rng(2)
    u = torch.tensor(rng.normal(size=(3, L.P)) * 0.5)
    series = torch.tensor([0, 1, 1])
    _, g = O.target_log_prob_and_grad(prob, u, series)
    h = 1e-6
    for i in range(L.P):
        up, um = u.clone(), u.clone()
        up[:, i] += h
        um[:, i] -= h
        fd = (O.target_log_prob(prob, up, series) - O.target_log_prob(prob, um, series)) / (2 * h)
        np.testing.assert_allclose(g[:, i].numpy(), fd.numpy(), rtol=1e-6, atol=1e-7)


def test_philox_matches_oracle_bit_for_bit():
    import ctypes
    from helpers import fptr, host_shim
    lib = host_shim()
    rng = np.random.default_rng(0)
    n = 257
    c = [rng.integers(0, 2 ** 32, size=n, dtype=np.uint32) for _ in range(4)]
    seed = 0x1234567890ABCDEF
    out = np.zeros((n, 4), dtype=np.uint32)
    lib.shim_philox4x32(n, fptr(c[0]), fptr(c[1]), fptr(c[2]), fptr(c[3]), ctypes.c_uint64(seed), fptr(out))
    ref = O.philox4x32(c[0], c[1], c[2], c[3], np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32))
    for i in range(4):
        np.testing.assert_array_equal(out[:, i], ref[i])
    # known-answer vector from the Random123 distribution (philox4x32-10, zero counter/key)
    z = O.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    nz = np.zeros((n, 4), dtype=np.float32)
    lib.shim_philox_normal4(n, fptr(c[0]), fptr(c[1]), fptr(c[2]), ctypes.c_uint32(7), ctypes.c_uint64(seed),
                            fptr(nz))
    refn = O.philox_normal4(c[0], c[1], c[2], 7, seed)
    for i in range(4):
        np.testing.assert_allclose(nz[:, i], refn[i], rtol=2e-5, atol=2e-6)
