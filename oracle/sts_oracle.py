"""CPU oracle for the tfcausalimpact fit/forecast hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``tfcausalimpact_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED at kernel level: the arithmetic of this path lives in the un-vendored
third-party dependency ``tensorflow-probability[tf] >=0.18,<=0.24`` (+ ``tensorflow>=2.10``),
pinned at /root/reference/setup.py:51-52 and absent from /root/reference and from this image
(no network).  The reference holds no golden vectors for log-prob / gradient / filter outputs
(/root/reference/tests/test_model.py:223-277 mocks TFP entirely).  This file therefore restates
TFP's *published* algorithm (tfp.sts.{Sum,LocalLevel,Seasonal,SparseLinearRegression},
tfd.LinearGaussianStateSpaceModel, tfp.sts.fit_with_hmc, tfp.vi.fit_surrogate_posterior,
tfp.sts.one_step_predictive, tfp.sts.forecast) following the reference's own call sites:

  * model recipe / priors ........ /root/reference/causalimpact/model.py:196-321
  * fit (hmc / vi) ............... /root/reference/causalimpact/model.py:356-386
  * predictive distributions ..... /root/reference/causalimpact/model.py:414-418, 446-451
  * sample reductions ............ /root/reference/causalimpact/inferences.py:52-249, 285-408
  * horseshoe weight formula ..... /root/reference/tests/test_main.py:326-339

What pins it: (1) the dense construction below builds the seasonal change-of-basis matrices
numerically (effects->residuals, permutation, residuals->effects) exactly the way TFP documents
them, independently of the closed forms the CUDA kernels use; (2) gradients come from torch
autograd and are cross-checked against fp64 central finite differences in tests/; (3) end-to-end
summaries are compared with the README / notebook / tests known answers (tests/golden/NOTEBOOK_TABLES.md,
tests/test_oracle_known_answers.py);
(4) the deterministic reduction semantics are checked against the reference's own
tests/test_inferences.py expectations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)
# wider model set (tfp.sts defaults a custom `model=` gets; /root/reference/causalimpact/main.py:186-252)
FLAG_DENSE_REG = 1        # tfp.sts.LinearRegression: weights ~ StudentT(5, 0, 10), identity bijector
FLAG_LLT = 2              # tfp.sts.LocalLinearTrend: level/slope scales ~ LogNormal(log(.05 sd), 3)
FLAG_OBS_LOGNORMAL = 4    # tfp.sts.Sum default observation_noise_scale prior LogNormal(log(.01 sd), 2)
FLAG_LEVEL_LOGNORMAL = 8  # tfp.sts.LocalLevel default level_scale prior LogNormal(log(.05 sd), 3)
# /root/reference/causalimpact/model.py:39
K_LOCAL_LEVEL_PRIOR_SAMPLE_SIZE = 32
# tfp.sts.SparseLinearRegression default `weights_prior_scale`; pinned by
# /root/reference/tests/test_main.py:329 (the `* 0.1`).
WEIGHTS_PRIOR_SCALE = 0.1


# ----------------------------------------------------------------------------------------
# Layout
# ----------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Layout:
    """Fixed state-space layout of Sum([LocalLevel, SparseLinearRegression?, Seasonal?]).

    Component order follows /root/reference/causalimpact/model.py:295,309,316.
    """
    T: int            # pre-period length (observed steps)
    Tf: int           # forecast horizon
    k: int            # number of covariates (0 => no regression component)
    S: int = 1        # nseasons (1 => no seasonal component)
    r: int = 1        # season_duration
    flags: int = 0    # wider model set (custom `model=`): FLAG_* below
    S2: int = 1       # num_seasons of a second tfp.sts.Seasonal component (1 => none); custom `model=` only
    r2: int = 1       # its num_steps_per_season

    @property
    def llt(self) -> bool:
        return bool(self.flags & FLAG_LLT)

    @property
    def dense_reg(self) -> bool:
        return bool(self.flags & FLAG_DENSE_REG)

    @property
    def nt(self) -> int:
        """trend latent size: level (+ slope for LocalLinearTrend)"""
        return 2 if self.llt else 1

    @property
    def d(self) -> int:
        return self.nt + (self.S - 1 if self.S > 1 else 0) + (self.S2 - 1 if self.S2 > 1 else 0)

    @property
    def nreg(self) -> int:
        return 0 if self.k == 0 else (self.k if self.dense_reg else 2 + 3 * self.k)

    @property
    def P(self) -> int:
        return 1 + self.nt + self.nreg + (1 if self.S > 1 else 0) + (1 if self.S2 > 1 else 0)

    # offsets into the flat parameter vector (TFP `model.parameters` order)
    @property
    def i_obs(self) -> int:
        return 0

    @property
    def i_level(self) -> int:
        return 1

    @property
    def i_slope(self) -> int:
        return 2

    @property
    def i_reg(self) -> int:
        return 1 + self.nt

    @property
    def i_gsv(self) -> int:
        return self.i_reg

    @property
    def i_gsn(self) -> int:
        return self.i_reg + 1

    @property
    def i_lsv(self) -> int:
        return self.i_reg + 2

    @property
    def i_lsn(self) -> int:
        return self.i_reg + 2 + self.k

    @property
    def i_wn(self) -> int:
        return self.i_reg + 2 + 2 * self.k

    @property
    def i_drift(self) -> int:
        return self.P - 1 - (1 if self.S2 > 1 else 0)

    @property
    def i_drift2(self) -> int:
        return self.P - 1

    def param_names(self) -> List[str]:
        names = ['observation_noise_scale', 'LocalLevel/_level_scale']
        if self.flags != 0 or self.S2 > 1:
            raise NotImplementedError('names of the wider model set live in tfcausalimpact_b200.model')
        if self.k > 0:
            names += ['SparseLinearRegression/_global_scale_variance',
                      'SparseLinearRegression/_global_scale_noncentered',
                      'SparseLinearRegression/_local_scale_variances',
                      'SparseLinearRegression/_local_scales_noncentered',
                      'SparseLinearRegression/_weights_noncentered']
        if self.S > 1:
            names += ['Seasonal/_drift_scale']
        return names

    def is_season_change(self, t: int) -> bool:
        """tfp.sts.Seasonal `is_last_day_of_season(t)` for scalar num_steps_per_season."""
        if self.S <= 1:
            return False
        return (t % (self.S * self.r)) % self.r == self.r - 1

    def is_season_change2(self, t: int) -> bool:
        if self.S2 <= 1:
            return False
        return (t % (self.S2 * self.r2)) % self.r2 == self.r2 - 1


@dataclass
class SeriesStats:
    """Per-series constants derived from the observed (pre-period) series.

    tfp.sts `empirical_statistics` (masked mean, population stddev, first unmasked value minus
    mean) + the reference's prior hyper-parameters (model.py:283-290).
    """
    mean: torch.Tensor      # [N]  constant_offset
    sd: torch.Tensor        # [N]  observed_stddev (ddof=0)
    y0c: torch.Tensor       # [N]  observed_initial (centered)
    obs_b: torch.Tensor     # [N]  InvGamma beta for observation noise variance
    level_b: torch.Tensor   # [N]  InvGamma beta for level variance
    drift_mu: torch.Tensor  # [N]  LogNormal loc for the seasonal drift scale


def series_stats(y: torch.Tensor, prior_level_sd: float = 0.01) -> SeriesStats:
    """y: [N, T] with NaN at missing steps."""
    mask = ~torch.isnan(y)
    cnt = mask.sum(dim=1).to(y.dtype)
    y0 = torch.where(mask, y, torch.zeros_like(y))
    mean = y0.sum(dim=1) / cnt
    var = (torch.where(mask, (y - mean[:, None]) ** 2, torch.zeros_like(y))).sum(dim=1) / cnt
    sd = torch.sqrt(var)
    first = mask.to(torch.int64).argmax(dim=1)
    y_first = y.gather(1, first[:, None])[:, 0]
    y0c = y_first - mean
    a = K_LOCAL_LEVEL_PRIOR_SAMPLE_SIZE / 2.0
    # model.py:289 obs prior uses sigma_guess = obs_sd / 2 with obs_sd = std(ddof=0) (model.py:283)
    obs_b = a * (sd / 2.0) ** 2
    level_b = torch.full_like(sd, a * prior_level_sd ** 2)
    drift_mu = torch.log(0.01 * sd)
    return SeriesStats(mean, sd, y0c, obs_b, level_b, drift_mu)


# ----------------------------------------------------------------------------------------
# Parameters: bijectors, priors, regression weights   
# ----------------------------------------------------------------------------------------
def softplus(u: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.softplus(u)


def log_sigmoid(u: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.logsigmoid(u)


def constrain(u: torch.Tensor, L: Layout, st: SeriesStats, series: torch.Tensor
              ) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """u: [B, P] unconstrained; series: [B] long index into the stats.

    Returns the constrained parameter dict (TFP names) and the forward log-det-Jacobian [B].
    Scale parameters use Scale(observed_stddev) o Softplus; horseshoe positives use Softplus;
    noncentered / dense weights are unconstrained.
    """
    sd = st.sd[series]
    th: Dict[str, torch.Tensor] = {}
    th['observation_noise_scale'] = sd * softplus(u[:, L.i_obs])
    th['LocalLevel/_level_scale'] = sd * softplus(u[:, L.i_level])
    ldj = 2.0 * torch.log(sd) + log_sigmoid(u[:, L.i_obs]) + log_sigmoid(u[:, L.i_level])
    if L.llt:
        th['LocalLinearTrend/_slope_scale'] = sd * softplus(u[:, L.i_slope])
        ldj = ldj + torch.log(sd) + log_sigmoid(u[:, L.i_slope])
    if L.k > 0 and L.dense_reg:
        th['LinearRegression/_weights'] = u[:, L.i_reg:L.i_reg + L.k]
    elif L.k > 0:
        th['SparseLinearRegression/_global_scale_variance'] = softplus(u[:, L.i_gsv])
        th['SparseLinearRegression/_global_scale_noncentered'] = softplus(u[:, L.i_gsn])
        th['SparseLinearRegression/_local_scale_variances'] = softplus(u[:, L.i_lsv:L.i_lsv + L.k])
        th['SparseLinearRegression/_local_scales_noncentered'] = softplus(u[:, L.i_lsn:L.i_lsn + L.k])
        th['SparseLinearRegression/_weights_noncentered'] = u[:, L.i_wn:L.i_wn + L.k]
        ldj = ldj + log_sigmoid(u[:, L.i_gsv]) + log_sigmoid(u[:, L.i_gsn]) \
            + log_sigmoid(u[:, L.i_lsv:L.i_lsv + L.k]).sum(dim=1) \
            + log_sigmoid(u[:, L.i_lsn:L.i_lsn + L.k]).sum(dim=1)
    if L.S > 1:
        th['Seasonal/_drift_scale'] = sd * softplus(u[:, L.i_drift])
        ldj = ldj + torch.log(sd) + log_sigmoid(u[:, L.i_drift])
    if L.S2 > 1:
        th['Seasonal2/_drift_scale'] = sd * softplus(u[:, L.i_drift2])
        ldj = ldj + torch.log(sd) + log_sigmoid(u[:, L.i_drift2])
    return th, ldj


def unconstrain(th: Dict[str, torch.Tensor], L: Layout, st: SeriesStats, series: torch.Tensor
                ) -> torch.Tensor:
    """Inverse of `constrain` (used to turn constrained draws back into kernel inputs)."""
    assert L.flags == 0 and L.S2 <= 1, 'default model only'

    def sp_inv(x):
        return x + torch.log(-torch.expm1(-x))
    sd = st.sd[series]
    cols = [sp_inv(th['observation_noise_scale'] / sd)[:, None],
            sp_inv(th['LocalLevel/_level_scale'] / sd)[:, None]]
    if L.k > 0:
        cols += [sp_inv(th['SparseLinearRegression/_global_scale_variance'])[:, None],
                 sp_inv(th['SparseLinearRegression/_global_scale_noncentered'])[:, None],
                 sp_inv(th['SparseLinearRegression/_local_scale_variances']),
                 sp_inv(th['SparseLinearRegression/_local_scales_noncentered']),
                 th['SparseLinearRegression/_weights_noncentered']]
    if L.S > 1:
        cols += [sp_inv(th['Seasonal/_drift_scale'] / sd)[:, None]]
    return torch.cat(cols, dim=1)


def _log_inv_gamma(x, a, b):
    # tfd.InverseGamma(concentration=a, scale=b).log_prob(x)
    a_t = torch.as_tensor(a, dtype=x.dtype)
    return a_t * torch.log(b) - torch.lgamma(a_t) - (a_t + 1.0) * torch.log(x) - b / x


def _log_lognormal(x, mu, sigma):
    z = (torch.log(x) - mu) / sigma
    return -torch.log(x) - math.log(sigma) - 0.5 * LOG_2PI - 0.5 * z * z


def log_prior(th: Dict[str, torch.Tensor], L: Layout, st: SeriesStats, series: torch.Tensor
              ) -> torch.Tensor:
    """Sum of prior log-densities at the constrained values.  [B]

    sigma = sqrt(V), V ~ InvGamma(16, b): TransformedDistribution(InverseGamma, Power(.5))
    (model.py:211-216, 234-235) => log p(sigma) = log IG(sigma^2) + log(2 sigma).  With the
    FLAG_* options the tfp.sts defaults of a custom model are used instead.
    """
    a = K_LOCAL_LEVEL_PRIOR_SAMPLE_SIZE / 2.0
    so = th['observation_noise_scale']
    sl = th['LocalLevel/_level_scale']
    mu01 = st.drift_mu[series]                       # log(0.01 sd)
    mu05 = mu01 + math.log(5.0)                      # log(0.05 sd)
    if L.flags & FLAG_OBS_LOGNORMAL:
        lp = _log_lognormal(so, mu01, 2.0)
    else:
        lp = _log_inv_gamma(so * so, a, st.obs_b[series]) + torch.log(2.0 * so)
    if L.llt or (L.flags & FLAG_LEVEL_LOGNORMAL):
        lp = lp + _log_lognormal(sl, mu05, 3.0)
    else:
        lp = lp + _log_inv_gamma(sl * sl, a, st.level_b[series]) + torch.log(2.0 * sl)
    if L.llt:
        lp = lp + _log_lognormal(th['LocalLinearTrend/_slope_scale'], mu05, 3.0)
    if L.k > 0 and L.dense_reg:
        w = th['LinearRegression/_weights']
        c = math.lgamma(3.0) - math.lgamma(2.5) - 0.5 * math.log(5.0 * math.pi) - math.log(10.0)
        lp = lp + (c - 3.0 * torch.log1p(w * w / 500.0)).sum(dim=1)
    elif L.k > 0:
        half = torch.as_tensor(0.5, dtype=so.dtype)
        gsv = th['SparseLinearRegression/_global_scale_variance']
        gsn = th['SparseLinearRegression/_global_scale_noncentered']
        lsv = th['SparseLinearRegression/_local_scale_variances']
        lsn = th['SparseLinearRegression/_local_scales_noncentered']
        wn = th['SparseLinearRegression/_weights_noncentered']
        log_hn = 0.5 * math.log(2.0 / math.pi)   # HalfNormal(1): log(2/sqrt(2 pi))
        lp = lp + _log_inv_gamma(gsv, 0.5, half) + (log_hn - 0.5 * gsn * gsn)
        lp = lp + _log_inv_gamma(lsv, 0.5, half).sum(dim=1) + (log_hn - 0.5 * lsn * lsn).sum(dim=1)
        lp = lp + (-0.5 * LOG_2PI - 0.5 * wn * wn).sum(dim=1)
    if L.S > 1:
        lp = lp + _log_lognormal(th['Seasonal/_drift_scale'], mu01, 3.0)
    if L.S2 > 1:
        lp = lp + _log_lognormal(th['Seasonal2/_drift_scale'], mu01, 3.0)
    return lp


def regression_weights(th: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Horseshoe weights, formula pinned by /root/reference/tests/test_main.py:326-339.  [B, k]
    (dense LinearRegression: the weights themselves)."""
    if 'LinearRegression/_weights' in th:
        return th['LinearRegression/_weights']
    global_scale = (th['SparseLinearRegression/_global_scale_noncentered'] *
                    torch.sqrt(th['SparseLinearRegression/_global_scale_variance']) *
                    WEIGHTS_PRIOR_SCALE)
    local_scales = (th['SparseLinearRegression/_local_scales_noncentered'] *
                    torch.sqrt(th['SparseLinearRegression/_local_scale_variances']))
    return th['SparseLinearRegression/_weights_noncentered'] * local_scales * global_scale[:, None]


# ----------------------------------------------------------------------------------------
# Dense TFP-style state space model -- built numerically
# ----------------------------------------------------------------------------------------
def _seasonal_basis(S: int, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """tfp.sts.Seasonal `build_effects_to_residuals_matrix`: (S-1)xS and Sx(S-1)."""
    full = np.eye(S) - 1.0 / S
    full[-1, :] = 1.0 / S
    inv = np.linalg.inv(full)
    return (torch.as_tensor(full[:-1, :], dtype=dtype), torch.as_tensor(inv[:, :-1], dtype=dtype))


class DenseSSM:
    """Time-varying dense matrices of the additive model for a batch of B parameter vectors.

    Latent = [level | seasonal residuals(S-1)].  (TFP's zero-size-effect dummy latent of the
    regression component -- transition 0, observation 0, prior N(0,0) -- is omitted: it
    contributes exact zeros to every product.)
    """

    def __init__(self, th: Dict[str, torch.Tensor], L: Layout, st: SeriesStats,
                 series: torch.Tensor):
        self.L = L
        dtype = th['observation_noise_scale'].dtype
        B = series.shape[0]
        d = L.d
        nt = L.nt
        self.B, self.d, self.dtype = B, d, dtype
        self.obs_var = th['observation_noise_scale'] ** 2
        self.level_var = th['LocalLevel/_level_scale'] ** 2
        H = torch.zeros(d, dtype=dtype)
        H[0] = 1.0
        F_base = torch.eye(d, dtype=dtype)
        self.Q_base = torch.zeros(B, d, d, dtype=dtype)
        self.Q_base[:, 0, 0] = self.level_var
        m0 = torch.zeros(B, d, dtype=dtype)
        m0[:, 0] = st.y0c[series]
        s0 = torch.abs(st.y0c[series]) + st.sd[series]
        P0 = torch.zeros(B, d, d, dtype=dtype)
        P0[:, 0, 0] = s0 * s0
        if L.llt:
            # tfp.sts.LocalLinearTrend: level' = level + slope + N(0, level_scale^2), slope' = slope +
            # N(0, slope_scale^2); initial slope ~ N(0, observed_stddev^2)
            F_base[0, 1] = 1.0
            self.Q_base[:, 1, 1] = th['LocalLinearTrend/_slope_scale'] ** 2
            P0[:, 1, 1] = st.sd[series] ** 2
        # seasonal blocks: (offset, S, A, Q block, change predicate)
        self.blocks = []
        off = nt
        for (S, name, pred) in ((L.S, 'Seasonal/_drift_scale', L.is_season_change),
                                (L.S2, 'Seasonal2/_drift_scale', L.is_season_change2)):
            if S <= 1:
                continue
            e2r, r2e = _seasonal_basis(S, dtype)
            perm = np.concatenate([np.arange(1, S), [0]])
            perm_m = torch.as_tensor(np.eye(S)[perm], dtype=dtype)
            A = e2r @ perm_m @ r2e                       # (S-1)x(S-1)
            h_eff = torch.zeros(1, S, dtype=dtype)
            h_eff[0, 0] = 1.0
            H[off:off + S - 1] = (h_eff @ r2e)[0]
            drift = th[name]
            Qb = ((drift / S) ** 2)[:, None, None] * torch.ones(S - 1, S - 1, dtype=dtype)
            # initial effects prior Normal(y0c, |y0c|+sd) per effect -> residual space
            mean_eff = st.y0c[series][:, None] * torch.ones(1, S, dtype=dtype)
            m0[:, off:off + S - 1] = mean_eff @ e2r.T
            cov_res = e2r @ e2r.T
            P0[:, off:off + S - 1, off:off + S - 1] = (s0 * s0)[:, None, None] * cov_res
            self.blocks.append((off, S, A, Qb, pred))
            off += S - 1
        self.H = H
        self.F_id = F_base
        self.m0, self.P0 = m0, P0
        self._cache = {}

    def _mats(self, t: int):
        key = tuple(bool(pred(t)) for (_, _, _, _, pred) in self.blocks)
        if key not in self._cache:
            F = self.F_id.clone()
            Q = self.Q_base.clone()
            for on, (off, S, A, Qb, _) in zip(key, self.blocks):
                if on:
                    F[off:off + S - 1, off:off + S - 1] = A
                    Q[:, off:off + S - 1, off:off + S - 1] = Qb
            self._cache[key] = (F, Q)
        return self._cache[key]

    def F(self, t: int) -> torch.Tensor:
        return self._mats(t)[0]

    def Q(self, t: int) -> torch.Tensor:
        return self._mats(t)[1]


def kalman_filter(ssm: DenseSSM, resid: torch.Tensor, t0: int = 0):
    """tfd.LinearGaussianStateSpaceModel.forward_filter for a scalar observation.

    resid: [B, T] = y - constant_offset - X beta (NaN = masked step).
    Joseph-form covariance update as TFP's `linear_gaussian_update`; masked steps skip the update
    and add 0 to the log-likelihood.
    Returns (loglik[B], innov_mean[B,T] (H m), innov_var[B,T] (s), m_pred[B,d], P_pred[B,d,d])
    where (m_pred, P_pred) is the predicted state after the last step (forecast prior).
    """
    B, T = resid.shape
    d = ssm.d
    m, P = ssm.m0, ssm.P0
    H = ssm.H
    eye = torch.eye(d, dtype=ssm.dtype)
    ll = torch.zeros(B, dtype=ssm.dtype)
    hm_list, s_list = [], []
    for t in range(T):
        r_t = resid[:, t]
        mask = torch.isnan(r_t)
        r_safe = torch.where(mask, torch.zeros_like(r_t), r_t)
        hm = m @ H
        g = P @ H                                       # [B, d]
        s = g @ H + ssm.obs_var
        v = r_safe - hm
        ll_t = -0.5 * (LOG_2PI + torch.log(s) + v * v / s)
        K = g / s[:, None]
        m_f = m + K * v[:, None]
        IKH = eye[None] - K[:, :, None] * H[None, None, :]
        P_f = IKH @ P @ IKH.transpose(1, 2) + ssm.obs_var[:, None, None] * (K[:, :, None] * K[:, None, :])
        ll = ll + torch.where(mask, torch.zeros_like(ll_t), ll_t)
        m_f = torch.where(mask[:, None], m, m_f)
        P_f = torch.where(mask[:, None, None], P, P_f)
        hm_list.append(hm)
        s_list.append(s)
        Ft = ssm.F(t0 + t)
        m = m_f @ Ft.T
        P = Ft[None] @ P_f @ Ft.T[None] + ssm.Q(t0 + t)
    return ll, torch.stack(hm_list, 1), torch.stack(s_list, 1), m, P


# ----------------------------------------------------------------------------------------
# Target log-prob (unconstrained) and gradient
# ----------------------------------------------------------------------------------------
@dataclass
class Problem:
    """A batch of N independent series sharing one layout."""
    L: Layout
    y: torch.Tensor            # [N, T]   (NaN = missing)
    X: Optional[torch.Tensor]  # [N, T+Tf, k]  or None when k == 0
    st: SeriesStats

    @staticmethod
    def build(y, X, L: Layout, prior_level_sd: float = 0.01, dtype=torch.float64) -> 'Problem':
        y = torch.as_tensor(np.asarray(y), dtype=dtype)
        Xt = None if L.k == 0 else torch.as_tensor(np.asarray(X), dtype=dtype)
        return Problem(L, y, Xt, series_stats(y, prior_level_sd))


def residuals(prob: Problem, th: Dict[str, torch.Tensor], series: torch.Tensor,
              t_lo: int, t_hi: int, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y_t - constant_offset - X_t beta for t in [t_lo, t_hi) (y omitted => minus-offset only)."""
    L = prob.L
    off = prob.st.mean[series][:, None].expand(-1, t_hi - t_lo)
    if L.k > 0:
        beta = regression_weights(th)                                   # [B, k]
        off = off + torch.einsum('btk,bk->bt', prob.X[series, t_lo:t_hi, :], beta)
    if y is None:
        return off
    return y - off


def joint_log_prob_parts(prob: Problem, u: torch.Tensor, series: torch.Tensor):
    th, ldj = constrain(u, prob.L, prob.st, series)
    lp = log_prior(th, prob.L, prob.st, series)
    ssm = DenseSSM(th, prob.L, prob.st, series)
    resid = residuals(prob, th, series, 0, prob.L.T, prob.y[series])
    ll, _, _, _, _ = kalman_filter(ssm, resid)
    return ll, lp, ldj


def target_log_prob(prob: Problem, u: torch.Tensor, series: torch.Tensor) -> torch.Tensor:
    """log p(y | theta) + log p(theta) + log|d theta / d u|  -- what HMC/VI see.  [B]"""
    ll, lp, ldj = joint_log_prob_parts(prob, u, series)
    return ll + lp + ldj


def target_log_prob_and_grad(prob: Problem, u: torch.Tensor, series: torch.Tensor):
    u = u.detach().clone().requires_grad_(True)
    lp = target_log_prob(prob, u, series)
    (g,) = torch.autograd.grad(lp.sum(), u)
    return lp.detach(), g.detach()


# ----------------------------------------------------------------------------------------
# Counter-based RNG shared with the device (Philox4x32-10 + Box-Muller)
# ----------------------------------------------------------------------------------------
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _PH_W0)
            k1 = np.uint32(k1 + _PH_W1)
    return c0, c1, c2, c3


def u32_to_unit(x):
    """(0,1) float32-exact uniform: ((x >> 8) + 0.5) * 2^-24."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) + 0.5) * (2.0 ** -24)


def philox_normal4(c0, c1, c2, c3, seed: int):
    """Four standard normals per counter (two Box-Muller pairs)."""
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    x0, x1, x2, x3 = philox4x32(c0, c1, c2, c3, k0, k1)
    u0, u1, u2, u3 = u32_to_unit(x0), u32_to_unit(x1), u32_to_unit(x2), u32_to_unit(x3)
    r0 = np.sqrt(-2.0 * np.log(u0))
    r1 = np.sqrt(-2.0 * np.log(u2))
    return (r0 * np.cos(2 * np.pi * u1), r0 * np.sin(2 * np.pi * u1),
            r1 * np.cos(2 * np.pi * u3), r1 * np.sin(2 * np.pi * u3))


def philox_uniform4(c0, c1, c2, c3, seed: int):
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    x0, x1, x2, x3 = philox4x32(c0, c1, c2, c3, k0, k1)
    return u32_to_unit(x0), u32_to_unit(x1), u32_to_unit(x2), u32_to_unit(x3)


# Stream tags (counter word 3) -- must match csrc/ci_rng.cuh
TAG_VI_INIT = 1
TAG_VI_EPS = 2
TAG_VI_DRAW = 3
TAG_HMC_MOM = 4
TAG_HMC_ACC = 5
TAG_ONESTEP = 6
TAG_FORECAST = 7
TAG_FORECAST_X0 = 8
TAG_PICK = 9


# ----------------------------------------------------------------------------------------
# VI  (model.py:366-386) and HMC (model.py:361-364)
# ----------------------------------------------------------------------------------------
def softplus_inv_scalar(x: float) -> float:
    return x + math.log(-math.expm1(-x))


def vi_fit(prob: Problem, series_of_unit: np.ndarray, num_steps: int, sample_size: int,
           seed: int, lr: float = 0.1, init_scale: float = 0.01,
           dtype=torch.float64) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    """Mean-field ADVI as tfp.vi.fit_surrogate_posterior runs it.

    One variational unit per entry of `series_of_unit` (series index).  Returns (mu[U,P],
    rho[U,P], loss[num_steps, U]).  q(u) = N(mu, softplus(rho)^2); loss = E_q[log q - log target]
    with `sample_size` reparameterised draws; Keras-legacy Adam(lr, .9, .999, 1e-7).
    """
    L = prob.L
    U = len(series_of_unit)
    Pn = L.P
    unit = np.arange(U, dtype=np.uint32)[:, None]
    pidx = np.arange(Pn, dtype=np.uint32)[None, :]
    un = philox_uniform4(unit, pidx, 0, TAG_VI_INIT, seed)[0]
    mu = torch.as_tensor(-2.0 + 4.0 * un, dtype=dtype)
    rho = torch.full((U, Pn), softplus_inv_scalar(init_scale), dtype=dtype)
    m_mu = torch.zeros_like(mu); v_mu = torch.zeros_like(mu)
    m_rho = torch.zeros_like(mu); v_rho = torch.zeros_like(mu)
    b1, b2, eps_adam = 0.9, 0.999, 1e-7
    series_lane = torch.as_tensor(np.repeat(series_of_unit, sample_size), dtype=torch.long)
    losses = np.zeros((num_steps, U))
    for step in range(num_steps):
        lane = (np.arange(U * sample_size, dtype=np.uint32))[:, None]
        xi = philox_normal4(lane, pidx, step, TAG_VI_EPS, seed)[0]       # [U*s, P]
        xi = torch.as_tensor(xi, dtype=dtype)
        sigma = softplus(rho)
        uu = (mu.repeat_interleave(sample_size, 0) + sigma.repeat_interleave(sample_size, 0) * xi)
        lp, g = target_log_prob_and_grad(prob, uu, series_lane)
        g = g.view(U, sample_size, Pn)
        xi_v = xi.view(U, sample_size, Pn)
        g_mu = -g.mean(dim=1)
        dsig = torch.sigmoid(rho)
        g_rho = (-(g * xi_v).mean(dim=1)) * dsig - dsig / sigma
        logq = (-torch.log(sigma)[:, None, :] - 0.5 * xi_v ** 2 - 0.5 * LOG_2PI).sum(dim=2)
        losses[step] = (logq - lp.view(U, sample_size)).mean(dim=1).numpy()
        t = step + 1
        lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        m_mu = b1 * m_mu + (1 - b1) * g_mu; v_mu = b2 * v_mu + (1 - b2) * g_mu * g_mu
        m_rho = b1 * m_rho + (1 - b1) * g_rho; v_rho = b2 * v_rho + (1 - b2) * g_rho * g_rho
        mu = mu - lr_t * m_mu / (torch.sqrt(v_mu) + eps_adam)
        rho = rho - lr_t * m_rho / (torch.sqrt(v_rho) + eps_adam)
    return mu, rho, losses


def vi_draw(mu: torch.Tensor, rho: torch.Tensor, num_draws: int, seed: int) -> torch.Tensor:
    """`surrogate.sample(num_draws)` in unconstrained space: [U, num_draws, P]."""
    U, Pn = mu.shape
    unit = np.arange(U, dtype=np.uint32)[:, None, None]
    dr = np.arange(num_draws, dtype=np.uint32)[None, :, None]
    pidx = np.arange(Pn, dtype=np.uint32)[None, None, :]
    xi = philox_normal4(unit * np.uint32(num_draws) + dr, pidx, 0, TAG_VI_DRAW, seed)[0]
    xi = torch.as_tensor(xi, dtype=mu.dtype)
    return mu[:, None, :] + softplus(rho)[:, None, :] * xi


def hmc_run(prob: Problem, series_of_chain: np.ndarray, u0: torch.Tensor, step0: torch.Tensor,
            num_results: int, num_warmup: int, num_leapfrog: int, seed: int,
            target_accept: float = 0.75, num_adapt: Optional[int] = None):
    """HMC + dual-averaging as tfp.sts.fit_with_hmc configures it.

    u0: [B, P] initial unconstrained state; step0: [B, P] per-coordinate initial step sizes.
    Returns (draws[num_results, B, P] unconstrained, accept[num_warmup+num_results, B],
    final_step_factor[B]).
    """
    B, Pn = u0.shape
    dtype = u0.dtype
    if num_adapt is None:
        num_adapt = int(num_warmup * 0.8)
    series = torch.as_tensor(series_of_chain, dtype=torch.long)
    lane = np.arange(B, dtype=np.uint32)[:, None]
    pidx = np.arange(Pn, dtype=np.uint32)[None, :]
    u = u0.clone()
    lp, g = target_log_prob_and_grad(prob, u, series)
    # dual averaging state (common multiplicative factor on all coordinates of a chain)
    log_fac = torch.zeros(B, dtype=dtype)        # eps = step0 * exp(log_fac)
    log_fac_avg = torch.zeros(B, dtype=dtype)
    err_sum = torch.zeros(B, dtype=dtype)
    gamma, t0, kappa = 0.05, 10.0, 0.75
    log10 = math.log(10.0)
    draws, accepts = [], []
    for it in range(num_warmup + num_results):
        eps = step0 * torch.exp(log_fac)[:, None]
        mom = torch.as_tensor(philox_normal4(lane, pidx, it, TAG_HMC_MOM, seed)[0], dtype=dtype)
        h0 = -lp + 0.5 * (mom * mom).sum(dim=1)
        uu, gg, pp = u.clone(), g.clone(), mom.clone()
        pp = pp + 0.5 * eps * gg
        lp_new = lp
        for l in range(num_leapfrog):
            uu = uu + eps * pp
            lp_new, gg = target_log_prob_and_grad(prob, uu, series)
            pp = pp + (1.0 if l < num_leapfrog - 1 else 0.5) * eps * gg
        h1 = -lp_new + 0.5 * (pp * pp).sum(dim=1)
        log_acc = h0 - h1
        log_acc = torch.where(torch.isfinite(log_acc), log_acc, torch.full_like(log_acc, -math.inf))
        un = torch.as_tensor(philox_uniform4(lane[:, 0], 0, it, TAG_HMC_ACC, seed)[0], dtype=dtype)
        acc = torch.log(un) < log_acc
        u = torch.where(acc[:, None], uu, u)
        g = torch.where(acc[:, None], gg, g)
        lp = torch.where(acc, lp_new, lp)
        accepts.append(acc.numpy().copy())
        if it < num_adapt:
            p_acc = torch.exp(torch.clamp(log_acc, max=0.0))
            err_sum = err_sum + target_accept - p_acc
            t = it + 1.0
            log_x = log10 - err_sum * math.sqrt(t) / ((t + t0) * gamma)
            eta = t ** (-kappa)
            log_fac_avg = eta * log_x + (1 - eta) * log_fac_avg
            log_fac = log_x if it < num_adapt - 1 else log_fac_avg
        if it >= num_warmup:
            draws.append(u.clone())
    return torch.stack(draws, 0), np.stack(accepts, 0), torch.exp(log_fac)


# ----------------------------------------------------------------------------------------
# Predictive distributions (model.py:414-418, 446-451)
# ----------------------------------------------------------------------------------------
def one_step_predictive(prob: Problem, u_draws: torch.Tensor, series: torch.Tensor):
    """means[D,T], variances[D,T] of the one-step-ahead predictive per posterior draw."""
    th, _ = constrain(u_draws, prob.L, prob.st, series)
    ssm = DenseSSM(th, prob.L, prob.st, series)
    off = residuals(prob, th, series, 0, prob.L.T)
    resid = prob.y[series] - off
    _, hm, s, m_pred, P_pred = kalman_filter(ssm, resid)
    return hm + off, s, m_pred, P_pred, th, ssm


def forecast_mean(prob: Problem, u_draws: torch.Tensor, series: torch.Tensor):
    """Per-draw forecast mean path [D, Tf] (mixture mean = average over draws)."""
    L = prob.L
    _, _, m, _, th, ssm = one_step_predictive(prob, u_draws, series)
    off = residuals(prob, th, series, L.T, L.T + L.Tf)
    out = []
    for t in range(L.Tf):
        out.append(m @ ssm.H + off[:, t])
        m = m @ ssm.F(L.T + t).T
    return torch.stack(out, 1)


def forecast_moments(prob: Problem, u_draws: torch.Tensor, series: torch.Tensor):
    """Per-draw forecast marginal means/variances [D,Tf] and the variance of the path sum [D].

    Used to check the Philox forecast sampler statistically (no RNG parity with TF is possible).
    """
    L = prob.L
    _, _, m, P, th, ssm = one_step_predictive(prob, u_draws, series)
    off = residuals(prob, th, series, L.T, L.T + L.Tf)
    means, vars_ = [], []
    for t in range(L.Tf):
        means.append(m @ ssm.H + off[:, t])
        vars_.append(torch.einsum('i,bij,j->b', ssm.H, P, ssm.H) + ssm.obs_var)
        Ft = ssm.F(L.T + t)
        m = m @ Ft.T
        P = Ft[None] @ P @ Ft.T[None] + ssm.Q(L.T + t)
    return torch.stack(means, 1), torch.stack(vars_, 1)


# ----------------------------------------------------------------------------------------
# Sample reductions (inferences.py) -- plain NumPy as the reference does
# ----------------------------------------------------------------------------------------
def percentiles(a: np.ndarray, alpha: float, axis: int = 0):
    """inferences.py:34-49 + np.percentile (linear interpolation)."""
    lo, hi = alpha * 100.0 / 2.0, 100 - alpha * 100.0 / 2.0
    return np.percentile(a, [lo, hi], axis=axis)


def reduce_inferences(y_pre, y_post, pre_mean, post_mean, sims_pre, sims_post, alpha):
    """The numeric content of compile_posterior_inferences (inferences.py:119-206).

    All inputs already in original units; y_post/post_* already masked.  Returns a dict of the 16
    columns as arrays: pre+post concatenations of length T+Tf', cumulative ones of length Tf'+1
    (zero-prepended, inferences.py:153-176, 186-206).
    """
    pre_lo, pre_hi = percentiles(sims_pre, alpha)
    post_lo, post_hi = percentiles(sims_post, alpha)
    cum_lo, cum_hi = percentiles(np.cumsum(sims_post, axis=1), alpha)
    ceff_lo, ceff_hi = percentiles(np.cumsum(y_post[None, :] - sims_post, axis=1), alpha)
    z = np.zeros(1)
    y_all = np.concatenate([y_pre, y_post])
    mean_all = np.concatenate([pre_mean, post_mean])
    lo_all = np.concatenate([pre_lo, post_lo])
    hi_all = np.concatenate([pre_hi, post_hi])
    return {
        'complete_preds_means': mean_all,
        'complete_preds_lower': lo_all,
        'complete_preds_upper': hi_all,
        'post_preds_means': post_mean,
        'post_preds_lower': post_lo,
        'post_preds_upper': post_hi,
        'post_cum_y': np.concatenate([z, np.cumsum(y_post)]),
        'post_cum_preds_means': np.concatenate([z, np.cumsum(post_mean)]),
        'post_cum_preds_lower': np.concatenate([z, cum_lo]),
        'post_cum_preds_upper': np.concatenate([z, cum_hi]),
        'point_effects_means': y_all - mean_all,
        'point_effects_lower': y_all - hi_all,
        'point_effects_upper': y_all - lo_all,
        'post_cum_effects_means': np.concatenate([z, np.cumsum(y_post - post_mean)]),
        'post_cum_effects_lower': np.concatenate([z, ceff_lo]),
        'post_cum_effects_upper': np.concatenate([z, ceff_hi]),
    }


def reduce_summary(y_post, post_mean, sims, alpha):
    """summarize_posterior_inferences (inferences.py:311-352) -> 10x2 array; rows in the
    reference's index order (inferences.py:355-367), columns (average, cumulative)."""
    mean_y, sum_y = y_post.mean(), y_post.sum()
    mean_p, sum_p = post_mean.mean(), post_mean.sum()
    mlo, mhi = percentiles(sims.mean(axis=1), alpha)
    slo, shi = percentiles(sims.sum(axis=1), alpha)
    rows = [
        [mean_y, sum_y], [mean_p, sum_p], [mlo, slo], [mhi, shi],
        [mean_y - mean_p, sum_y - sum_p], [mean_y - mhi, sum_y - shi], [mean_y - mlo, sum_y - slo],
        [(mean_y - mean_p) / mean_p, (sum_y - sum_p) / sum_p],
        [(mean_y - mhi) / mean_p, (sum_y - shi) / sum_p],
        [(mean_y - mlo) / mean_p, (sum_y - slo) / sum_p],
    ]
    return np.asarray(rows, dtype=np.float64)


def p_value(sims, post_sum):
    """compute_p_value (inferences.py:397-408): strict inequalities, /(n+1)."""
    s = sims.sum(axis=1)
    return min(np.sum(s > post_sum), np.sum(s < post_sum)) / (len(sims) + 1)


# ----------------------------------------------------------------------------------------
# Component decomposition (tfp.sts.decompose_by_component / decompose_forecast_by_component, used
# by the reference's notebook: notebooks/getting_started.ipynb:125,144) -- Rauch-Tung-Striebel
# smoother over the dense model, per posterior draw, then the mixture over draws.
# ----------------------------------------------------------------------------------------
def _mixture_moments(mean_jt: torch.Tensor, var_jt: torch.Tensor, G: int):
    """[N*G, W] per-draw moments -> mixture mean / stddev [N, W]."""
    NG, W = mean_jt.shape
    m = mean_jt.view(NG // G, G, W)
    v = var_jt.view(NG // G, G, W)
    mix_mean = m.mean(dim=1)
    mix_var = (v + m * m).mean(dim=1) - mix_mean * mix_mean
    return mix_mean, torch.sqrt(torch.clamp(mix_var, min=0.0))


def smooth_components(prob: Problem, u_draws: torch.Tensor, series: torch.Tensor, G: int):
    """Smoothed component marginals.  Returns dict name -> (mean[N,T], stddev[N,T]) with names
    'LocalLevel', 'SparseLinearRegression' (if k > 0), 'Seasonal' (if S > 1)."""
    L = prob.L
    th, _ = constrain(u_draws, L, prob.st, series)
    ssm = DenseSSM(th, L, prob.st, series)
    off = residuals(prob, th, series, 0, L.T)
    resid = prob.y[series] - off
    B, T = resid.shape
    d = ssm.d
    eye = torch.eye(d, dtype=ssm.dtype)
    m, P = ssm.m0, ssm.P0
    mp, Pp, mf, Pf = [], [], [], []
    for t in range(T):
        mp.append(m); Pp.append(P)
        r_t = resid[:, t]
        mask = torch.isnan(r_t)
        r_safe = torch.where(mask, torch.zeros_like(r_t), r_t)
        g = P @ ssm.H
        s = g @ ssm.H + ssm.obs_var
        v = r_safe - m @ ssm.H
        K = g / s[:, None]
        m_f = torch.where(mask[:, None], m, m + K * v[:, None])
        P_f = torch.where(mask[:, None, None], P, P - K[:, :, None] * g[:, None, :])
        mf.append(m_f); Pf.append(P_f)
        Ft = ssm.F(t)
        m = m_f @ Ft.T
        P = Ft[None] @ P_f @ Ft.T[None] + ssm.Q(t)
    ms, Ps = [None] * T, [None] * T
    ms[T - 1], Ps[T - 1] = mf[T - 1], Pf[T - 1]
    for t in range(T - 2, -1, -1):
        Ft = ssm.F(t)
        A = Pf[t] @ Ft.T[None]                                  # P_f F^T
        J = torch.linalg.solve(Pp[t + 1], A.transpose(1, 2)).transpose(1, 2)   # A P_pred^-1 (P_pred symmetric)
        ms[t] = mf[t] + (J @ (ms[t + 1] - mp[t + 1])[:, :, None])[:, :, 0]
        Ps[t] = Pf[t] + J @ (Ps[t + 1] - Pp[t + 1]) @ J.transpose(1, 2)
    ms = torch.stack(ms, 1)                                     # [B, T, d]
    Ps = torch.stack(Ps, 1)                                     # [B, T, d, d]
    out = {'LocalLevel': _mixture_moments(ms[:, :, 0], Ps[:, :, 0, 0], G)}   # trend: the level entry
    if L.k > 0:
        reg = off - prob.st.mean[series][:, None]
        out['SparseLinearRegression'] = _mixture_moments(reg, torch.zeros_like(reg), G)
    # a Seasonal component contributes its current effect = first residual of its block (H picks it)
    if L.S > 1:
        i1 = L.nt
        out['Seasonal'] = _mixture_moments(ms[:, :, i1], Ps[:, :, i1, i1], G)
    if L.S2 > 1:
        i2 = L.nt + L.S - 1
        out['Seasonal2'] = _mixture_moments(ms[:, :, i2], Ps[:, :, i2, i2], G)
    return out


def forecast_components(prob: Problem, u_draws: torch.Tensor, series: torch.Tensor, G: int):
    """Per-component forecast marginals (prior propagation from the filtered state at T)."""
    L = prob.L
    _, _, m, P, th, ssm = one_step_predictive(prob, u_draws, series)
    off = residuals(prob, th, series, L.T, L.T + L.Tf)
    lm, lv, sm, sv, sm2, sv2 = [], [], [], [], [], []
    i1, i2 = L.nt, L.nt + L.S - 1
    for t in range(L.Tf):
        lm.append(m[:, 0]); lv.append(P[:, 0, 0])
        if L.S > 1:
            sm.append(m[:, i1]); sv.append(P[:, i1, i1])
        if L.S2 > 1:
            sm2.append(m[:, i2]); sv2.append(P[:, i2, i2])
        Ft = ssm.F(L.T + t)
        m = m @ Ft.T
        P = Ft[None] @ P @ Ft.T[None] + ssm.Q(L.T + t)
    out = {'LocalLevel': _mixture_moments(torch.stack(lm, 1), torch.stack(lv, 1), G)}
    if L.k > 0:
        reg = off - prob.st.mean[series][:, None]
        out['SparseLinearRegression'] = _mixture_moments(reg, torch.zeros_like(reg), G)
    if L.S > 1:
        out['Seasonal'] = _mixture_moments(torch.stack(sm, 1), torch.stack(sv, 1), G)
    if L.S2 > 1:
        out['Seasonal2'] = _mixture_moments(torch.stack(sm2, 1), torch.stack(sv2, 1), G)
    return out
