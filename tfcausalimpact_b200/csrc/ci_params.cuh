// Parameter block of the fixed layout: unconstrained vector u[P] -> constrained values, regression
// weights beta, prior log-density + bijector log-det, and the matching vector-Jacobian product.
//
// Parameter order (tfp.sts.Sum `model.parameters`, component order from
// /root/reference/causalimpact/model.py:295,309,316):
//   0 observation_noise_scale      sd * softplus(u)   sqrt(V), V ~ InvGamma(16, 16 (sd_y/2)^2)   model.py:289-290
//   1 LocalLevel/_level_scale      sd * softplus(u)   sqrt(V), V ~ InvGamma(16, 16 prior_level_sd^2) model.py:284-285
//   [k > 0]  2 global_scale_variance (softplus, InvGamma(.5,.5)), 3 global_scale_noncentered
//            (softplus, HalfNormal(1)), 4..4+k local_scale_variances, 4+k..4+2k
//            local_scales_noncentered, 4+2k..4+3k weights_noncentered (identity, Normal(0,1))
//   [S > 1]  P-1 Seasonal/_drift_scale  sd * softplus(u)  LogNormal(log(.01 sd), 3)
// beta = wn * (lsn sqrt(lsv)) * (gsn sqrt(gsv) 0.1)   -- /root/reference/tests/test_main.py:326-339
#pragma once
#include "ci_core.cuh"

#ifndef CI_PB
#define CI_PB 4   // covariates per load group in the parameter transforms
#endif

namespace ci {

// Per-series constants, struct-of-arrays on the device: sconst[field * N + n].
enum SConstField : int {
  SC_MEAN = 0,       // constant_offset = mean(y_pre)
  SC_SD = 1,         // observed_stddev (ddof = 0)
  SC_M0 = 2,         // initial level mean = y_first - mean
  SC_P0 = 3,         // (|y0c| + sd)^2
  SC_OBS_B = 4,      // InvGamma scale for sigma_obs^2
  SC_LEVEL_B = 5,    // InvGamma scale for sigma_level^2
  SC_DRIFT_MU = 6,   // log(0.01 sd)
  SC_PRIOR_CONST = 7,  // theta-independent additive constant of log prior + log-det
  SC_NFIELDS = 8
};

#define CI_PRIOR_A 16.0f          // kLocalLevelPriorSampleSize / 2   (model.py:39, 211-213)
#define CI_WEIGHTS_PRIOR_SCALE 0.1f

CI_HD int num_params(int k, int S) { return 2 + (k > 0 ? 2 + 3 * k : 0) + (S > 1 ? 1 : 0); }

struct LaneScalars {
  float R, ql, qd, lp0;
};

// u, beta: element stride `ld`.  sc: pointer to this series' constants with field stride `scld`.
CI_HD LaneScalars params_forward_lane(const float* u, long ld, int k, int S,
                                      const float* __restrict__ sc, long scld, float* __restrict__ beta,
                                      long ldb, float beta_sign = 1.0f) {
  const float sd = sc[SC_SD * scld];
  LaneScalars o;
  const float u0 = u[0], u1 = u[ld];
  const SoftplusPack p0 = softplus_pack(u0), p1 = softplus_pack(u1);
  const float so = sd * p0.sp, sl = sd * p1.sp;
  const float so2 = so * so, sl2 = sl * sl;
  float lp = sc[SC_PRIOR_CONST * scld];
  lp += -(2.f * CI_PRIOR_A + 1.f) * logf(so) - sc[SC_OBS_B * scld] / so2 + p0.lsig;
  lp += -(2.f * CI_PRIOR_A + 1.f) * logf(sl) - sc[SC_LEVEL_B * scld] / sl2 + p1.lsig;
  o.R = so2;
  o.ql = sl2;
  o.qd = 0.f;
  if (k > 0) {
    const float ug = u[2 * ld], un = u[3 * ld];
    const SoftplusPack pg = softplus_pack(ug), pn = softplus_pack(un);
    const float gsv = pg.sp, gsn = pn.sp;
    lp += -1.5f * logf(gsv) - 0.5f / gsv - 0.5f * gsn * gsn + pg.lsig + pn.lsig;
    const float G = gsn * sqrtf(gsv) * CI_WEIGHTS_PRIOR_SCALE * beta_sign;
    const float* ulsv = u + 4 * ld;
    const float* ulsn = u + (long)(4 + k) * ld;
    const float* uwn = u + (long)(4 + 2 * k) * ld;
    // covariates in groups of CI_PB: the 3 CI_PB loads of the next group are in flight while the current
    // group is processed, so the (DRAM) latency of the parameter columns is paid k / CI_PB times, not k times
    float a[CI_PB], b[CI_PB], wn[CI_PB];
#pragma unroll
    for (int q = 0; q < CI_PB; ++q) {
      const long o = (long)(q < k ? q : k - 1) * ld;
      a[q] = ulsv[o]; b[q] = ulsn[o]; wn[q] = uwn[o];
    }
    for (int j0 = 0; j0 < k; j0 += CI_PB) {
      float an[CI_PB], bn[CI_PB], wnn[CI_PB];
#pragma unroll
      for (int q = 0; q < CI_PB; ++q) {
        const int jn = j0 + CI_PB + q;
        const long o = (long)(jn < k ? jn : k - 1) * ld;
        an[q] = ulsv[o]; bn[q] = ulsn[o]; wnn[q] = uwn[o];
      }
#pragma unroll
      for (int q = 0; q < CI_PB; ++q) {
        if (j0 + q < k) {
          const SoftplusPack pa = softplus_pack(a[q]), pb = softplus_pack(b[q]);
          const float lsv = pa.sp, lsn = pb.sp;
          lp += -1.5f * logf(lsv) - 0.5f * rcp(lsv) - 0.5f * lsn * lsn - 0.5f * wn[q] * wn[q] + pa.lsig + pb.lsig;
          beta[(long)(j0 + q) * ldb] = wn[q] * lsn * sqrtf(lsv) * G;
        }
      }
#pragma unroll
      for (int q = 0; q < CI_PB; ++q) {
        a[q] = an[q]; b[q] = bn[q]; wn[q] = wnn[q];
      }
    }
  }
  if (S > 1) {
    const float ud = u[(long)(num_params(k, S) - 1) * ld];
    const SoftplusPack pd = softplus_pack(ud);
    const float sdr = sd * pd.sp;
    const float lg = logf(sdr);
    const float z = (lg - sc[SC_DRIFT_MU * scld]) * (1.0f / 3.0f);
    lp += -lg - 0.5f * z * z + pd.lsig;
    const float q = sdr * (1.0f / S);
    o.qd = q * q;
  }
  o.lp0 = lp;
  return o;
}

// grad[i] = d/du_i [ log prior + log-det + log-lik ], given the filter adjoints
// (Rb, qlb, qdb) and beta_bar = dloglik/dbeta (stride ldb).
CI_HD void params_backward_lane(const float* u, long ld, int k, int S,
                                const float* __restrict__ sc, long scld,
                                const float* __restrict__ beta_bar, long ldb, float Rb, float qlb, float qdb,
                                float* __restrict__ grad) {
  const float sd = sc[SC_SD * scld];
  {
    const float u0 = u[0];
    const SoftplusPack p = softplus_pack(u0);
    const float so = sd * p.sp;
    const float dth = -(2.f * CI_PRIOR_A + 1.f) / so + 2.f * sc[SC_OBS_B * scld] / (so * so * so) + 2.f * so * Rb;
    grad[0] = dth * sd * p.sig + p.nsig;
  }
  {
    const float u1 = u[ld];
    const SoftplusPack p = softplus_pack(u1);
    const float sl = sd * p.sp;
    const float dth = -(2.f * CI_PRIOR_A + 1.f) / sl + 2.f * sc[SC_LEVEL_B * scld] / (sl * sl * sl) + 2.f * sl * qlb;
    grad[ld] = dth * sd * p.sig + p.nsig;
  }
  if (k > 0) {
    const float ug = u[2 * ld], un = u[3 * ld];
    const SoftplusPack pg = softplus_pack(ug), pn = softplus_pack(un);
    const float gsv = pg.sp, gsn = pn.sp;
    const float sq_gsv = sqrtf(gsv);
    const float G = gsn * sq_gsv * CI_WEIGHTS_PRIOR_SCALE;
    const float* ulsv = u + 4 * ld;
    const float* ulsn = u + (long)(4 + k) * ld;
    const float* uwn = u + (long)(4 + 2 * k) * ld;
    float* glsv = grad + 4 * ld;
    float* glsn = grad + (long)(4 + k) * ld;
    float* gwn = grad + (long)(4 + 2 * k) * ld;
    float dG = 0.f;
    float a[CI_PB], b[CI_PB], wn[CI_PB];
#pragma unroll
    for (int q = 0; q < CI_PB; ++q) {
      const long o = (long)(q < k ? q : k - 1) * ld;
      a[q] = ulsv[o]; b[q] = ulsn[o]; wn[q] = uwn[o];
    }
    for (int j0 = 0; j0 < k; j0 += CI_PB) {
      float an[CI_PB], bn[CI_PB], wnn[CI_PB];
#pragma unroll
      for (int q = 0; q < CI_PB; ++q) {
        const int jn = j0 + CI_PB + q;
        const long o = (long)(jn < k ? jn : k - 1) * ld;
        an[q] = ulsv[o]; bn[q] = ulsn[o]; wnn[q] = uwn[o];
      }
#pragma unroll
      for (int q = 0; q < CI_PB; ++q) {
        const int j = j0 + q;
        if (j < k) {
          const float bb = beta_bar[(long)j * ldb];
          const SoftplusPack pa = softplus_pack(a[q]), pb = softplus_pack(b[q]);
          const float lsv = pa.sp, lsn = pb.sp;
          const float sq = sqrtf(lsv);
          const float ilsv = rcp(lsv);
          const float l = lsn * sq;
          dG = fmaf(bb, wn[q] * l, dG);
          gwn[(long)j * ld] = bb * l * G - wn[q];
          const float dl = bb * wn[q] * G;
          glsn[(long)j * ld] = (dl * sq - lsn) * pb.sig + pb.nsig;
          // d sqrt(v)/dv = sqrt(v) / (2 v);  d log IG(.5,.5)/dv = -1.5 / v + 0.5 / v^2
          glsv[(long)j * ld] = (dl * lsn * 0.5f * sq * ilsv + ilsv * (0.5f * ilsv - 1.5f)) * pa.sig + pa.nsig;
        }
      }
#pragma unroll
      for (int q = 0; q < CI_PB; ++q) {
        a[q] = an[q]; b[q] = bn[q]; wn[q] = wnn[q];
      }
    }
    grad[3 * ld] = (dG * sq_gsv * CI_WEIGHTS_PRIOR_SCALE - gsn) * pn.sig + pn.nsig;
    grad[2 * ld] = (dG * gsn * CI_WEIGHTS_PRIOR_SCALE * 0.5f / sq_gsv - 1.5f / gsv + 0.5f / (gsv * gsv)) * pg.sig +
                   pg.nsig;
  }
  if (S > 1) {
    const long id = (long)(num_params(k, S) - 1) * ld;
    const float ud = u[id];
    const SoftplusPack pd = softplus_pack(ud);
    const float sdr = sd * pd.sp;
    const float lg = logf(sdr);
    const float dth = -1.0f / sdr - (lg - sc[SC_DRIFT_MU * scld]) / (9.0f * sdr) + 2.f * sdr * qdb * (1.0f / (S * S));
    grad[id] = dth * sd * pd.sig + pd.nsig;
  }
}

// Constrained values theta[i] (same order/stride) from u -- what `model_samples` exposes.
CI_HD void params_constrain_lane(const float* __restrict__ u, long ld, int k, int S, float sd,
                                 float* __restrict__ theta, long ldo) {
  theta[0] = sd * softplus(u[0]);
  theta[ldo] = sd * softplus(u[ld]);
  if (k > 0) {
    for (int i = 2; i < 4 + 2 * k; ++i) theta[(long)i * ldo] = softplus(u[(long)i * ld]);
    for (int i = 4 + 2 * k; i < 4 + 3 * k; ++i) theta[(long)i * ldo] = u[(long)i * ld];
  }
  if (S > 1) {
    const long i = num_params(k, S) - 1;
    theta[i * ldo] = sd * softplus(u[i * ld]);
  }
}

}  // namespace ci
