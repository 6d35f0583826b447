"""Device-side engine: owns the HBM-resident inputs of a batch of series and drives the C-ABI.

Layout in HBM (float32, lane = fastest axis of every per-lane array):
    y      [N][T]          observed series (NaN = missing)
    XY     [Tpad/4][N][k+1][4] chunk-major covariates + y (one TMA bulk copy per warp and chunk)
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
                 device=None, flags=0, nseasons2=0, season_duration2=0):
        nt.require_cuda()
        self.lib = nt.lib()
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        y = torch.as_tensor(np.array(y, dtype=np.float32) if not torch.is_tensor(y) else y)
        if y.dim() == 1:
            y = y[None, :]
        self.y = y.to(self.device, torch.float32).contiguous()
        self.N, self.T = self.y.shape
        self.Tf, self.S, self.r = int(Tf), int(nseasons), int(season_duration)
        self.flags = int(flags)      # wider model set (_native.FLAG_*): runs on the warp-per-lane kernels
        # second Seasonal component of a custom model (Sum([..., Seasonal(S), Seasonal(S2)])); 0 = none
        self.S2, self.r2 = (int(nseasons2), int(season_duration2) or 1) if int(nseasons2) > 1 else (0, 0)
        if X is None:
            self.k = 0
        else:
            X = torch.as_tensor(np.array(X, dtype=np.float32) if not torch.is_tensor(X) else X)
            if X.dim() == 2:
                X = X[None]
            assert X.shape[0] == self.N and X.shape[1] == self.T + self.Tf, \
                f'design matrix must be [N, T+Tf, k], got {tuple(X.shape)}'
            self.k = X.shape[2]
        self.Tpad = (self.T + self.Tf + 3) // 4 * 4
        self.P = nt.num_params(self.k, self.S, self.flags, self.S2)
        self.sconst = torch.empty((8, self.N), dtype=torch.float32, device=self.device)
        L = self.layout(1)
        with torch.cuda.device(self.device):
            nt.check(self.lib.ci_series_stats(ctypes.byref(L), nt.ptr(self.y), ctypes.c_float(prior_level_sd),
                                              nt.ptr(self.sconst), nt.stream_ptr()), 'ci_series_stats')
            Xrm = X.to(self.device, torch.float32).contiguous() if self.k > 0 else None
            self.X = torch.empty((self.Tpad // 4, self.N, self.k + 1, 4), dtype=torch.float32, device=self.device)
            nt.check(self.lib.ci_pack_design(ctypes.byref(L), nt.ptr(Xrm), nt.ptr(self.y), nt.ptr(self.sconst), nt.ptr(self.X),
                                             nt.stream_ptr()), 'ci_pack_design')
            # covariate-major copy [Tpad/4][k+1][N][4] for calls with fewer than 4 lanes per series (one chain /
            # one VI sample, the reference's defaults): built on first use, see data()
            self.XT = None
            if Xrm is not None:
                # stream-ordered: Xrm may be released once the pack kernel has been enqueued
                Xrm.record_stream(torch.cuda.current_stream())
        self._ws = {}

    # ------------------------------------------------------------------------------------------
    def layout(self, G):
        return nt.CiLayout(self.N, int(G), self.T, self.Tf, self.k, self.S, self.r, self.Tpad, self.flags, self.S2, self.r2)

    def data(self, G=None):
        """ci_data for a call with G lanes per series; G < 4 adds the covariate-major design block."""
        xt = None
        if G is not None and G < 4 and self.k >= 1 and self.lib.ci_eval_uses_xt(ctypes.byref(self.layout(G))):
            if self.XT is None:
                L = self.layout(1)
                self.XT = torch.empty((self.Tpad // 4, self.k + 1, self.N, 4), dtype=torch.float32, device=self.device)
                with torch.cuda.device(self.device):
                    nt.check(self.lib.ci_repack_design_cov(ctypes.byref(L), nt.ptr(self.X), nt.ptr(self.XT),
                                                           nt.stream_ptr()), 'ci_repack_design_cov')
            xt = self.XT
        return nt.CiData(nt.ptr(self.y).value, nt.ptr(self.X).value, nt.ptr(self.sconst).value,
                         nt.ptr(xt).value if xt is not None else None)

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
        D = self.data(B // self.N)
        with torch.cuda.device(self.device):
            nt.check(self.lib.ci_kalman_logp_grad(ctypes.byref(L), ctypes.byref(D), nt.ptr(u.contiguous()),
                                                  nt.ptr(logp), nt.ptr(grad), nt.ptr(ws),
                                                  ctypes.c_size_t(ws.numel()), nt.stream_ptr()),
                     'ci_kalman_logp_grad')
        return logp, grad

    # ------------------------------------------------------------------------------------------
    def _call(self, name, *args):
        with torch.cuda.device(self.device):
            nt.check(getattr(self.lib, name)(*args), name)

    def _f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def vi_fit(self, units_per_series=1, sample_size=1, num_steps=200, learning_rate=0.1, init_scale=0.01,
               seed=0, return_loss=False):
        """Mean-field ADVI (model.py:367-381).  Returns mu[P,U], rho[P,U] (+ loss[num_steps,U])."""
        U = self.N * units_per_series
        L = self.layout(units_per_series * sample_size)
        mu, rho = self._f32(self.P, U), self._f32(self.P, U)
        loss = self._f32(num_steps, U) if return_loss else None
        ws = self.workspace('fit', self.lib.ci_vi_workspace_bytes(ctypes.byref(L), ctypes.c_int32(units_per_series)))
        D = self.data(units_per_series * sample_size)
        self._call('ci_vi_fit', ctypes.byref(L), ctypes.byref(D), ctypes.c_int32(units_per_series),
                   ctypes.c_int32(sample_size), ctypes.c_int32(num_steps), ctypes.c_float(learning_rate),
                   ctypes.c_float(init_scale), ctypes.c_uint64(seed), nt.ptr(mu), nt.ptr(rho), nt.ptr(loss),
                   nt.ptr(ws), ctypes.c_size_t(ws.numel()), nt.stream_ptr())
        return (mu, rho, loss) if return_loss else (mu, rho)

    def vi_draw(self, mu, rho, num_draws, seed=0):
        """`surrogate_posterior.sample(num_draws)` in unconstrained space: [P, U * num_draws]."""
        U = mu.shape[1]
        out = self._f32(self.P, U * num_draws)
        self._call('ci_vi_draw', ctypes.c_int32(self.P), ctypes.c_int64(U), ctypes.c_int32(num_draws),
                   ctypes.c_uint64(seed), nt.ptr(mu), nt.ptr(rho), nt.ptr(out), nt.stream_ptr())
        return out

    def hmc_run(self, u, step0, num_results=100, num_warmup=50, num_adapt=None, num_leapfrog=15,
                target_accept=0.75, seed=0, iter_offset=0, draws=None):
        """HMC + dual averaging (model.py:361-364).  u [P,B] is updated in place.
        Returns draws [num_results, P, B] (unconstrained) and stats [4, B]."""
        assert u.is_cuda and u.is_contiguous() and u.shape[0] == self.P
        B = u.shape[1]
        L = self.layout(B // self.N)
        if num_adapt is None:
            num_adapt = int(num_warmup * 0.8)
        if draws is None:
            draws = self._f32(max(num_results, 1), self.P, B)
        stats = self._f32(4, B)
        ws = self.workspace('fit', self.lib.ci_hmc_workspace_bytes(ctypes.byref(L)))
        D = self.data(B // self.N)
        self._call('ci_hmc_run', ctypes.byref(L), ctypes.byref(D), nt.ptr(u), nt.ptr(step0.contiguous()),
                   ctypes.c_int32(num_results), ctypes.c_int32(num_warmup), ctypes.c_int32(num_adapt),
                   ctypes.c_int32(num_leapfrog), ctypes.c_float(target_accept), ctypes.c_uint64(seed),
                   ctypes.c_int32(iter_offset), nt.ptr(draws), nt.ptr(stats), nt.ptr(ws),
                   ctypes.c_size_t(ws.numel()), nt.stream_ptr())
        return draws[:num_results], stats

    def constrain(self, u):
        """Constrained parameter values theta [P,B] (what `model_samples` exposes)."""
        B = u.shape[1]
        L = self.layout(B // self.N)
        theta = torch.empty_like(u)
        D = self.data()
        self._call('ci_constrain', ctypes.byref(L), ctypes.byref(D), nt.ptr(u.contiguous()), nt.ptr(theta),
                   nt.stream_ptr())
        return theta

    # ---- predictive path ---------------------------------------------------------------------
    def filter_predict(self, u):
        """One-step-ahead predictive mean/variance [T,B] per draw + forecast prior state."""
        B = u.shape[1]
        L = self.layout(B // self.N)
        nfs = int(self.lib.ci_fstate_floats(ctypes.byref(L)))
        pm, pv, fs = self._f32(self.T, B), self._f32(self.T, B), self._f32(nfs, B)
        ws = self.workspace('pred', self.lib.ci_predict_workspace_bytes(ctypes.byref(L)))
        D = self.data()
        self._call('ci_filter_predict', ctypes.byref(L), ctypes.byref(D), nt.ptr(u.contiguous()), nt.ptr(pm),
                   nt.ptr(pv), nt.ptr(fs), nt.ptr(ws), ctypes.c_size_t(ws.numel()), nt.stream_ptr())
        return pm, pv, fs

    def onestep_sample(self, pm, pv, nsamp, seed=0, mu_sig=None, want_mean=True):
        B = pm.shape[1]
        L = self.layout(B // self.N)
        sims = self._f32(self.N, nsamp, self.T) if nsamp > 0 else None
        mean = self._f32(self.N, self.T) if want_mean else None
        self._call('ci_onestep_sample', ctypes.byref(L), nt.ptr(pm), nt.ptr(pv), nt.ptr(mu_sig),
                   ctypes.c_int32(nsamp), ctypes.c_uint64(seed), nt.ptr(sims), nt.ptr(mean), nt.stream_ptr())
        return sims, mean

    def forecast_sample(self, u, fs, nsamp, seed=0, mu_sig=None, want_mean=True):
        B = u.shape[1]
        L = self.layout(B // self.N)
        sims = self._f32(self.N, nsamp, self.Tf) if nsamp > 0 else None
        mean = self._f32(self.N, self.Tf) if want_mean else None
        ws = self.workspace('pred', self.lib.ci_predict_workspace_bytes(ctypes.byref(L)))
        D = self.data()
        self._call('ci_forecast_sample', ctypes.byref(L), ctypes.byref(D), nt.ptr(u.contiguous()), nt.ptr(fs),
                   nt.ptr(mu_sig), ctypes.c_int32(nsamp), ctypes.c_uint64(seed), nt.ptr(sims), nt.ptr(mean),
                   nt.ptr(ws), ctypes.c_size_t(ws.numel()), nt.stream_ptr())
        return sims, mean

    # ---- component decomposition -------------------------------------------------------------
    def decompose(self, u, fstate=None):
        """Mixture-over-draws component marginals [C, N, W] (mean, sd): smoothed over the observed
        period (fstate=None, W = T) or forecast over the horizon (fstate given, W = Tf).
        Rows: trend, regression, seasonal (, second seasonal: C = 4 when the model has one, else 3)."""
        B = u.shape[1]
        L = self.layout(B // self.N)
        W = self.T if fstate is None else self.Tf
        C = 4 if self.S2 > 1 else 3
        mean, sd = self._f32(C, self.N, W), self._f32(C, self.N, W)
        nbytes = self.lib.ci_decompose_workspace_bytes(ctypes.byref(L))
        if nbytes == 0:
            raise nt.NativeError(f'component decomposition does not support nseasons={self.S}')
        ws = self.workspace('dec', nbytes)
        D = self.data()
        if fstate is None:
            self._call('ci_decompose_by_component', ctypes.byref(L), ctypes.byref(D), nt.ptr(u.contiguous()),
                       nt.ptr(mean), nt.ptr(sd), nt.ptr(ws), ctypes.c_size_t(ws.numel()), nt.stream_ptr())
        else:
            self._call('ci_decompose_forecast', ctypes.byref(L), ctypes.byref(D), nt.ptr(u.contiguous()),
                       nt.ptr(fstate), nt.ptr(mean), nt.ptr(sd), nt.ptr(ws), ctypes.c_size_t(ws.numel()),
                       nt.stream_ptr())
        return mean, sd

    # ---- reductions --------------------------------------------------------------------------
    def reduce_inferences(self, y_pre, y_post, pre_sims, pre_mean, post_sims, post_mean, alpha):
        return reduce_inferences(y_pre, y_post, pre_sims, pre_mean, post_sims, post_mean, alpha)

    def reduce_summary(self, y_post, post_mean, post_sims, alpha):
        return reduce_summary(y_post, post_mean, post_sims, alpha)

    # ---- host-side conveniences (plumbing, not on the hot path) ---------------------------------
    def set_obs_prior_scale(self, scale):
        """Override the InverseGamma scale of the observation-noise variance prior (the device
        default is 16 (sd_y / 2)^2, model.py:289-290) and patch the prior constant accordingly."""
        new = torch.full((self.N,), float(scale), dtype=torch.float32, device=self.device)
        old = self.sconst[4].clone()
        if torch.allclose(old, new, rtol=1e-5, atol=0.0):
            return
        self.sconst[7] += 16.0 * (torch.log(new) - torch.log(old))
        self.sconst[4] = new

    def unconstrain(self, theta):
        """Inverse of `constrain`: theta [P,B] -> u [P,B] (torch ops; used once per fit)."""
        B = theta.shape[1]
        G = B // self.N
        sd = self.sconst[1].repeat_interleave(G)[None, :]

        def sp_inv(x):
            return x + torch.log(-torch.expm1(-x))
        u = theta.clone()
        k = self.k
        nscale = 1 + (2 if self.flags & nt.FLAG_LLT else 1)
        u[0:nscale] = sp_inv(theta[0:nscale] / sd)
        if k > 0 and not (self.flags & nt.FLAG_DENSE_REG):
            u[nscale:nscale + 2 + 2 * k] = sp_inv(theta[nscale:nscale + 2 + 2 * k])
        nseas = (1 if self.S > 1 else 0) + (1 if self.S2 > 1 else 0)
        if nseas:
            u[self.P - nseas:] = sp_inv(theta[self.P - nseas:] / sd)
        return u.contiguous()


def _layout_for(N, T, Tf):
    return nt.CiLayout(int(N), 1, int(T), int(Tf), 0, 1, 1, (int(T) + int(Tf) + 3) // 4 * 4, 0, 0, 0)


def reduce_inferences(y_pre, y_post, pre_sims, pre_mean, post_sims, post_mean, alpha):
    """Device reduction of compile_posterior_inferences (inferences.py:119-206).
    y_pre [N,T], y_post [N,Tf] (NaN = masked), pre_sims [N,n,T], post_sims [N,n,Tf] -> [N,16,T+Tf+1]."""
    nt.require_cuda()
    lib = nt.lib()
    N, T = y_pre.shape
    Tf, nsamp = y_post.shape[1], pre_sims.shape[1]
    L = _layout_for(N, T, Tf)
    dev = y_pre.device
    out = torch.empty((N, 16, T + Tf + 1), dtype=torch.float32, device=dev)
    nbytes = lib.ci_reduce_workspace_bytes(ctypes.byref(L), ctypes.c_int32(nsamp))
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        nt.check(lib.ci_reduce_inferences(ctypes.byref(L), nt.ptr(y_pre), nt.ptr(y_post), nt.ptr(pre_sims),
                                          nt.ptr(pre_mean), nt.ptr(post_sims), nt.ptr(post_mean),
                                          ctypes.c_int32(nsamp), ctypes.c_float(alpha), nt.ptr(out), nt.ptr(ws),
                                          ctypes.c_size_t(ws.numel()), nt.stream_ptr()), 'ci_reduce_inferences')
    return out


def reduce_summary(y_post, post_mean, post_sims, alpha):
    """Device reduction of summarize_posterior_inferences + compute_p_value -> [N,21]."""
    nt.require_cuda()
    lib = nt.lib()
    N, Tf = y_post.shape
    nsamp = post_sims.shape[1]
    L = _layout_for(N, 1, Tf)
    out = torch.empty((N, 21), dtype=torch.float32, device=y_post.device)
    with torch.cuda.device(y_post.device):
        nt.check(lib.ci_reduce_summary(ctypes.byref(L), nt.ptr(y_post), nt.ptr(post_mean), nt.ptr(post_sims),
                                       ctypes.c_int32(nsamp), ctypes.c_float(alpha), nt.ptr(out),
                                       ctypes.c_void_p(0), ctypes.c_size_t(0), nt.stream_ptr()), 'ci_reduce_summary')
    return out


class HostUploader:
    """Double-buffered host -> device staging on a side stream, so the PCIe copy of batch i+1
    overlaps the kernels of batch i (pinned host tensors in, device tensors + ready event out)."""

    def __init__(self, device):
        nt.require_cuda()
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def start(self, *host_tensors):
        """Enqueue the copies on the side stream; returns (device tensors, event)."""
        with torch.cuda.stream(self.stream):
            out = [None if t is None else t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    @staticmethod
    def finish(tensors, ev):
        """Make the current stream wait for the copies and adopt the tensors."""
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for t in tensors:
            if t is not None:
                t.record_stream(cur)
        return tensors
