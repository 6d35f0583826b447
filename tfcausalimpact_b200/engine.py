"""Device-side engine: owns the HBM-resident inputs of a batch of series and drives the C-ABI.

Layout in HBM (float32, lane = fastest axis of every per-lane array):
    y      [N][T]          observed series (NaN = missing)
    X      [N][k][Tpad]    packed design matrix, time contiguous (16-byte loads of 4 steps)
    sconst [8][N]          per-series constants (mean, sd, m0, P0, prior hyper-parameters)
    u/grad [P][B]          unconstrained parameters / gradient, B = N * G lanes
"""
import ctypes

import numpy as np
import torch

from . import _native as nt


class DeviceBatch:
    """N independent series sharing one state-space layout, resident on one GPU."""

    def __init__(self, y, X=None, Tf=0, nseasons=1, season_duration=1, prior_level_sd=0.01,
                 device=None):
        nt.require_cuda()
        self.lib = nt.lib()
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        y = torch.as_tensor(np.asarray(y, dtype=np.float32) if not torch.is_tensor(y) else y)
        if y.dim() == 1:
            y = y[None, :]
        self.y = y.to(self.device, torch.float32).contiguous()
        self.N, self.T = self.y.shape
        self.Tf, self.S, self.r = int(Tf), int(nseasons), int(season_duration)
        if X is None:
            self.k = 0
        else:
            X = torch.as_tensor(np.asarray(X, dtype=np.float32) if not torch.is_tensor(X) else X)
            if X.dim() == 2:
                X = X[None]
            assert X.shape[0] == self.N and X.shape[1] == self.T + self.Tf, \
                f'design matrix must be [N, T+Tf, k], got {tuple(X.shape)}'
            self.k = X.shape[2]
        self.Tpad = (self.T + self.Tf + 3) // 4 * 4
        self.P = nt.num_params(self.k, self.S)
        self.sconst = torch.empty((8, self.N), dtype=torch.float32, device=self.device)
        L = self.layout(1)
        with torch.cuda.device(self.device):
            nt.check(self.lib.ci_series_stats(ctypes.byref(L), nt.ptr(self.y), ctypes.c_float(prior_level_sd),
                                              nt.ptr(self.sconst), nt.stream_ptr()), 'ci_series_stats')
            if self.k > 0:
                Xrm = X.to(self.device, torch.float32).contiguous()
                self.X = torch.empty((self.N, self.k, self.Tpad), dtype=torch.float32, device=self.device)
                nt.check(self.lib.ci_pack_design(ctypes.byref(L), nt.ptr(Xrm), nt.ptr(self.X), nt.stream_ptr()),
                         'ci_pack_design')
                # stream-ordered: Xrm may be released once the pack kernel has been enqueued
                Xrm.record_stream(torch.cuda.current_stream())
            else:
                self.X = None
        self._ws = {}

    # ------------------------------------------------------------------------------------------
    def layout(self, G):
        return nt.CiLayout(self.N, int(G), self.T, self.Tf, self.k, self.S, self.r, self.Tpad)

    def data(self):
        return nt.CiData(nt.ptr(self.y).value, nt.ptr(self.X).value if self.X is not None else None,
                         nt.ptr(self.sconst).value)

    def workspace(self, kind, nbytes):
        buf = self._ws.get(kind)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[kind] = buf
        return buf

    # ------------------------------------------------------------------------------------------
    def logp_grad(self, u, logp=None, grad=None):
        """Joint log-prob and gradient in unconstrained space.  u: [P, B] CUDA float32."""
        assert u.is_cuda and u.dtype == torch.float32 and u.dim() == 2 and u.shape[0] == self.P
        B = u.shape[1]
        assert B % self.N == 0, 'lanes must be a multiple of the number of series'
        L = self.layout(B // self.N)
        if logp is None:
            logp = torch.empty(B, dtype=torch.float32, device=self.device)
        if grad is None:
            grad = torch.empty_like(u)
        nbytes = self.lib.ci_eval_workspace_bytes(ctypes.byref(L))
        ws = self.workspace('eval', nbytes)
        D = self.data()
        with torch.cuda.device(self.device):
            nt.check(self.lib.ci_kalman_logp_grad(ctypes.byref(L), ctypes.byref(D), nt.ptr(u.contiguous()),
                                                  nt.ptr(logp), nt.ptr(grad), nt.ptr(ws),
                                                  ctypes.c_size_t(ws.numel()), nt.stream_ptr()),
                     'ci_kalman_logp_grad')
        return logp, grad
