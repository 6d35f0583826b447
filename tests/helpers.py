"""Shared test helpers: synthetic problems, packing into the device layout, oracle glue."""
import ctypes
import math
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sts_oracle as O  # noqa: E402

SC_NFIELDS = 8


def synth_problem(N, T, Tf, k, S, r=1, seed=0, nan_frac=0.0, standardize=True):
    """Synthetic series per SURVEY 8(d): level RW + seasonal + sparse regression + noise."""
    rng = np.random.default_rng(seed)
    Tt = T + Tf
    X = rng.normal(size=(N, Tt, max(k, 1))).astype(np.float32)[:, :, :k]
    beta = np.zeros((N, k))
    nz = int(math.ceil(k / 5)) if k > 0 else 0
    if nz:
        beta[:, :nz] = rng.normal(size=(N, nz))
    level = np.cumsum(rng.normal(scale=0.05, size=(N, Tt)), axis=1)
    y = level + rng.normal(scale=0.3, size=(N, Tt))
    if S > 1:
        eff = rng.normal(scale=0.5, size=(N, S))
        eff -= eff.mean(axis=1, keepdims=True)
        idx = (np.arange(Tt) // r) % S
        y = y + eff[:, idx]
    if k > 0:
        y = y + np.einsum('ntk,nk->nt', X, beta)
    y[:, T:] += 0.5
    if standardize:
        mu = y[:, :T].mean(axis=1, keepdims=True)
        sd = y[:, :T].std(axis=1, keepdims=True)
        y = (y - mu) / sd
    y = y.astype(np.float32)
    if nan_frac > 0:
        m = rng.random(size=(N, T)) < nan_frac
        m[:, 0] = False
        y[:, :T][m] = np.nan
    return y[:, :T].copy(), y[:, T:].copy(), X


def pack_X(X, y, Tpad=None):
    """X [N, Ttot, k], y [N, T] -> chunk-major XY [Tpad/4, N, k+1, 4] (row k = y - mean(y), NaN past T)."""
    N, Tt, k = X.shape
    T = y.shape[1]
    if Tpad is None:
        Tpad = (Tt + 3) // 4 * 4
    buf = np.zeros((N, Tpad, k + 1), dtype=np.float32)
    buf[:, :Tt, :k] = X
    buf[:, :, k] = np.nan
    buf[:, :T, k] = y - np.nanmean(y.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    out = np.ascontiguousarray(buf.reshape(N, Tpad // 4, 4, k + 1).transpose(1, 0, 3, 2))
    return out, Tpad


def prior_const(L: 'O.Layout', obs_b, level_b, sd):
    a = O.K_LOCAL_LEVEL_PRIOR_SAMPLE_SIZE / 2.0
    c = (a * math.log(obs_b) - math.lgamma(a)) + (a * math.log(level_b) - math.lgamma(a))
    c += 2 * math.log(2.0) + 2 * math.log(sd)
    if L.k > 0:
        c += (1 + L.k) * (0.5 * math.log(0.5) - math.lgamma(0.5))
        c += (1 + L.k) * 0.5 * math.log(2.0 / math.pi)
        c += -L.k * 0.5 * O.LOG_2PI
    if L.S > 1:
        c += -math.log(3.0) - 0.5 * O.LOG_2PI + math.log(sd)
    return c


def sconst_from_oracle(prob: 'O.Problem'):
    """Device per-series constant table [8, N] from the oracle's fp64 statistics."""
    st, L = prob.st, prob.L
    N = prob.y.shape[0]
    sc = np.zeros((SC_NFIELDS, N), dtype=np.float64)
    sc[0] = st.mean.numpy()
    sc[1] = st.sd.numpy()
    sc[2] = st.y0c.numpy()
    sc[3] = (np.abs(st.y0c.numpy()) + st.sd.numpy()) ** 2
    sc[4] = st.obs_b.numpy()
    sc[5] = st.level_b.numpy()
    sc[6] = st.drift_mu.numpy()
    for n in range(N):
        sc[7, n] = prior_const(L, sc[4, n], sc[5, n], sc[1, n])
    return sc.astype(np.float32)


_SHIM = None


def host_shim():
    """Build (if needed) and load the TEST-ONLY host shim of the per-lane kernel arithmetic."""
    global _SHIM
    if _SHIM is not None:
        return _SHIM
    d = os.path.join(ROOT, 'tests', 'host_shim')
    so = os.path.join(d, 'libci_hostshim.so')
    src = os.path.join(d, 'shim.cpp')
    inc = os.path.join(ROOT, 'tfcausalimpact_b200', 'csrc')
    newest = max(os.path.getmtime(os.path.join(inc, f)) for f in os.listdir(inc) if f.endswith('.cuh'))
    newest = max(newest, os.path.getmtime(src))
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-x', 'c++', '-std=c++17', '-I', inc, src,
                               '-o', so])
    _SHIM = ctypes.CDLL(so)
    return _SHIM


def fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def shim_logp_grad(L, XY, Tpad, sconst, u_pb, G):
    """XY: chunk-major [Tpad/4, N, k+1, 4]; u_pb: [P, B] float32 (lane fastest).
    Returns logp[B], grad[P, B]."""
    lib = host_shim()
    N = XY.shape[1]
    B = N * G
    logp = np.zeros(B, dtype=np.float32)
    grad = np.zeros((L.P, B), dtype=np.float32)
    XY = np.ascontiguousarray(XY, dtype=np.float32)
    sconst = np.ascontiguousarray(sconst, dtype=np.float32)
    u_pb = np.ascontiguousarray(u_pb, dtype=np.float32)
    rc = lib.shim_logp_grad(N, G, L.T, L.k, L.S, L.r, fptr(XY), Tpad, fptr(sconst), fptr(u_pb), fptr(logp),
                            fptr(grad))
    assert rc == 0, rc
    return logp, grad
