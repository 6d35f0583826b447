// Parameter block of the WIDER model set, used by the large-latent kernels when
// ci_layout.flags != 0:
//   trend      LocalLevel (1 scale) | LocalLinearTrend (level + slope scales)        CI_FLAG_LLT
//   regression SparseLinearRegression (horseshoe, 2 + 3k) | LinearRegression (k weights) CI_FLAG_DENSE_REG
//   priors     observation noise: sqrt-InverseGamma (tfcausalimpact default model, model.py:289-290)
//              | LogNormal(log(.01 sd), 2) (tfp.sts.Sum default, custom `model=`)   CI_FLAG_OBS_LOGNORMAL
//              level scale: sqrt-InverseGamma (model.py:284-285) | LogNormal(log(.05 sd), 3)
//              (tfp.sts.LocalLevel / LocalLinearTrend default)                      CI_FLAG_LEVEL_LOGNORMAL / LLT
//              slope scale: LogNormal(log(.05 sd), 3); dense weights: StudentT(5, 0, 10) [TFP defaults]
//   seasonal   up to TWO tfp.sts.Seasonal components (S, r) and (S2, r2), each with its own drift scale
//              ~ LogNormal(log(.01 sd), 3) (notebooks/getting_started.ipynb cell 116: Month + Year)
// Parameter order = tfp `model.parameters`: observation_noise_scale, trend scales, regression, drift(s).
// Every constant of the log-density is evaluated explicitly here (this path is not hot), so
// SC_PRIOR_CONST is not used.  These are the custom-model examples of the reference's docs
// (/root/reference/causalimpact/main.py:186-252, notebooks/getting_started.ipynb cell 44).
#pragma once
#include "ci_core.cuh"
#include "ci_params.cuh"

#define CI_FLAG_DENSE_REG 1
#define CI_FLAG_LLT 2
#define CI_FLAG_OBS_LOGNORMAL 4
#define CI_FLAG_LEVEL_LOGNORMAL 8

namespace ci {

CI_HD int num_params_gen(int k, int S, int flags, int S2 = 1) {
  const int trend = (flags & CI_FLAG_LLT) ? 2 : 1;
  const int reg = k > 0 ? ((flags & CI_FLAG_DENSE_REG) ? k : 2 + 3 * k) : 0;
  return 1 + trend + reg + (S > 1 ? 1 : 0) + (S2 > 1 ? 1 : 0);
}

struct GenScalars {
  float R, ql, qs, qd, lp0, qd2;
};

// log density of theta = sd softplus(u) under the scale prior + log |d theta / d u|, and its
// derivative with respect to u given dL/d(theta^2) (= `vbar`, the filter adjoint of the variance).
struct ScaleTerm {
  float theta, lp, dlp_du;
};
CI_HD ScaleTerm scale_term(float u, float sd, bool lognormal, float ig_b, float ln_mu, float ln_sigma, float vbar) {
  const SoftplusPack p = softplus_pack(u);
  ScaleTerm o;
  o.theta = sd * p.sp;
  const float lt = logf(o.theta);
  float dth;
  if (lognormal) {
    const float z = (lt - ln_mu) / ln_sigma;
    o.lp = -lt - logf(ln_sigma) - 0.5f * CI_LOG_2PI - 0.5f * z * z;
    dth = -1.0f / o.theta - z / (ln_sigma * o.theta);
  } else {
    // theta = sqrt(V), V ~ InvGamma(a, b): log IG(theta^2) + log(2 theta)
    const float a = CI_PRIOR_A;
    o.lp = a * logf(ig_b) - lgammaf(a) - (2.f * a + 1.f) * lt - ig_b / (o.theta * o.theta) + 0.6931471805599453f;
    dth = -(2.f * a + 1.f) / o.theta + 2.f * ig_b / (o.theta * o.theta * o.theta);
  }
  o.lp += logf(sd) + p.lsig;                       // bijector log-det
  dth += 2.f * o.theta * vbar;                     // through theta^2
  o.dlp_du = dth * sd * p.sig + p.nsig;
  return o;
}

// Forward: returns variances + log prior/log-det, writes beta (stride ldb).
CI_HD GenScalars params_forward_gen(const float* __restrict__ u, long ld, int k, int S, int flags,
                                    const float* __restrict__ sc, long scld, float* __restrict__ beta, long ldb,
                                    int S2 = 1) {
  const float sd = sc[SC_SD * scld];
  const float mu01 = sc[SC_DRIFT_MU * scld];                 // log(0.01 sd)
  const float mu05 = mu01 + 1.6094379124341003f;             // log(0.05 sd)
  const bool llt = flags & CI_FLAG_LLT;
  GenScalars o;
  int i = 0;
  const ScaleTerm so = scale_term(u[(long)i++ * ld], sd, flags & CI_FLAG_OBS_LOGNORMAL, sc[SC_OBS_B * scld], mu01, 2.f, 0.f);
  const ScaleTerm sl = scale_term(u[(long)i++ * ld], sd, llt || (flags & CI_FLAG_LEVEL_LOGNORMAL), sc[SC_LEVEL_B * scld],
                                  mu05, 3.f, 0.f);
  float lp = so.lp + sl.lp;
  o.R = so.theta * so.theta;
  o.ql = sl.theta * sl.theta;
  o.qs = 0.f;
  o.qd = 0.f;
  o.qd2 = 0.f;
  if (llt) {
    const ScaleTerm ss = scale_term(u[(long)i++ * ld], sd, true, 0.f, mu05, 3.f, 0.f);
    lp += ss.lp;
    o.qs = ss.theta * ss.theta;
  }
  if (k > 0) {
    if (flags & CI_FLAG_DENSE_REG) {
      // StudentT(df = 5, loc = 0, scale = 10): lgamma(3) - lgamma(2.5) - .5 log(5 pi) - log 10 - 3 log(1 + w^2 / 500)
      const float c = 0.6931471805599453f - 0.2846828704729192f - 0.5f * logf(5.f * 3.14159265358979f) - 2.302585092994046f;
      for (int j = 0; j < k; ++j) {
        const float w = u[(long)(i + j) * ld];
        lp += c - 3.f * log1pf(w * w * (1.0f / 500.f));
        beta[(long)j * ldb] = w;
      }
      i += k;
    } else {
      const float ug = u[(long)i * ld], un = u[(long)(i + 1) * ld];
      const SoftplusPack pg = softplus_pack(ug), pn = softplus_pack(un);
      const float lig = 0.5f * logf(0.5f) - lgammaf(0.5f);           // InvGamma(.5, .5) constant
      const float lhn = 0.5f * logf(2.f / 3.14159265358979f);         // HalfNormal(1) constant
      lp += lig - 1.5f * logf(pg.sp) - 0.5f / pg.sp + pg.lsig + lhn - 0.5f * pn.sp * pn.sp + pn.lsig;
      const float G = pn.sp * sqrtf(pg.sp) * CI_WEIGHTS_PRIOR_SCALE;
      for (int j = 0; j < k; ++j) {
        const float a = u[(long)(i + 2 + j) * ld], b = u[(long)(i + 2 + k + j) * ld], wn = u[(long)(i + 2 + 2 * k + j) * ld];
        const SoftplusPack pa = softplus_pack(a), pb = softplus_pack(b);
        lp += lig - 1.5f * logf(pa.sp) - 0.5f / pa.sp + pa.lsig + lhn - 0.5f * pb.sp * pb.sp + pb.lsig -
              0.5f * CI_LOG_2PI - 0.5f * wn * wn;
        beta[(long)j * ldb] = wn * pb.sp * sqrtf(pa.sp) * G;
      }
      i += 2 + 3 * k;
    }
  }
  if (S > 1) {
    const ScaleTerm sdft = scale_term(u[(long)i * ld], sd, true, 0.f, mu01, 3.f, 0.f);
    lp += sdft.lp;
    const float q = sdft.theta * (1.0f / S);
    o.qd = q * q;
    ++i;
  }
  if (S2 > 1) {
    const ScaleTerm sdft = scale_term(u[(long)i * ld], sd, true, 0.f, mu01, 3.f, 0.f);
    lp += sdft.lp;
    const float q = sdft.theta * (1.0f / S2);
    o.qd2 = q * q;
  }
  o.lp0 = lp;
  return o;
}

// Backward: grad[i] = d/du_i [log prior + log-det + log-lik] given the filter adjoints of the
// variances (Rb, qlb, qsb, qdb) and beta_bar (stride ldb).
CI_HD void params_backward_gen(const float* __restrict__ u, long ld, int k, int S, int flags,
                               const float* __restrict__ sc, long scld, const float* __restrict__ beta_bar, long ldb,
                               float Rb, float qlb, float qsb, float qdb, float* __restrict__ grad, int S2 = 1,
                               float qd2b = 0.f) {
  const float sd = sc[SC_SD * scld];
  const float mu01 = sc[SC_DRIFT_MU * scld];
  const float mu05 = mu01 + 1.6094379124341003f;
  const bool llt = flags & CI_FLAG_LLT;
  int i = 0;
  grad[(long)i * ld] = scale_term(u[(long)i * ld], sd, flags & CI_FLAG_OBS_LOGNORMAL, sc[SC_OBS_B * scld], mu01, 2.f, Rb).dlp_du;
  ++i;
  grad[(long)i * ld] = scale_term(u[(long)i * ld], sd, llt || (flags & CI_FLAG_LEVEL_LOGNORMAL), sc[SC_LEVEL_B * scld], mu05,
                                  3.f, qlb).dlp_du;
  ++i;
  if (llt) {
    grad[(long)i * ld] = scale_term(u[(long)i * ld], sd, true, 0.f, mu05, 3.f, qsb).dlp_du;
    ++i;
  }
  if (k > 0) {
    if (flags & CI_FLAG_DENSE_REG) {
      for (int j = 0; j < k; ++j) {
        const float w = u[(long)(i + j) * ld];
        grad[(long)(i + j) * ld] = beta_bar[(long)j * ldb] - 6.f * w / (500.f + w * w);
      }
      i += k;
    } else {
      const float ug = u[(long)i * ld], un = u[(long)(i + 1) * ld];
      const SoftplusPack pg = softplus_pack(ug), pn = softplus_pack(un);
      const float gsv = pg.sp, gsn = pn.sp, sq_gsv = sqrtf(gsv);
      const float G = gsn * sq_gsv * CI_WEIGHTS_PRIOR_SCALE;
      float dG = 0.f;
      for (int j = 0; j < k; ++j) {
        const long ia = (long)(i + 2 + j) * ld, ib = (long)(i + 2 + k + j) * ld, iw = (long)(i + 2 + 2 * k + j) * ld;
        const float a = u[ia], b = u[ib], wn = u[iw], bb = beta_bar[(long)j * ldb];
        const SoftplusPack pa = softplus_pack(a), pb = softplus_pack(b);
        const float lsv = pa.sp, lsn = pb.sp, sq = sqrtf(lsv), ilsv = 1.0f / lsv;
        const float l = lsn * sq;
        dG = fmaf(bb, wn * l, dG);
        grad[iw] = bb * l * G - wn;
        const float dl = bb * wn * G;
        grad[ib] = (dl * sq - lsn) * pb.sig + pb.nsig;
        grad[ia] = (dl * lsn * 0.5f * sq * ilsv + ilsv * (0.5f * ilsv - 1.5f)) * pa.sig + pa.nsig;
      }
      grad[(long)(i + 1) * ld] = (dG * sq_gsv * CI_WEIGHTS_PRIOR_SCALE - gsn) * pn.sig + pn.nsig;
      grad[(long)i * ld] = (dG * gsn * CI_WEIGHTS_PRIOR_SCALE * 0.5f / sq_gsv - 1.5f / gsv + 0.5f / (gsv * gsv)) * pg.sig +
                           pg.nsig;
      i += 2 + 3 * k;
    }
  }
  if (S > 1) {
    grad[(long)i * ld] = scale_term(u[(long)i * ld], sd, true, 0.f, mu01, 3.f, qdb * (1.0f / (S * S))).dlp_du;
    ++i;
  }
  if (S2 > 1) grad[(long)i * ld] = scale_term(u[(long)i * ld], sd, true, 0.f, mu01, 3.f, qd2b * (1.0f / (S2 * S2))).dlp_du;
}

// Constrained values theta (same order / stride) from u.
CI_HD void params_constrain_gen(const float* __restrict__ u, long ld, int k, int S, int flags, float sd,
                                float* __restrict__ theta, long ldo, int S2 = 1) {
  const int nscale = 1 + ((flags & CI_FLAG_LLT) ? 2 : 1);
  int i = 0;
  for (; i < nscale; ++i) theta[(long)i * ldo] = sd * softplus(u[(long)i * ld]);
  if (k > 0) {
    if (flags & CI_FLAG_DENSE_REG) {
      for (int j = 0; j < k; ++j, ++i) theta[(long)i * ldo] = u[(long)i * ld];
    } else {
      for (int j = 0; j < 2 + 2 * k; ++j, ++i) theta[(long)i * ldo] = softplus(u[(long)i * ld]);
      for (int j = 0; j < k; ++j, ++i) theta[(long)i * ldo] = u[(long)i * ld];
    }
  }
  if (S > 1) {
    theta[(long)i * ldo] = sd * softplus(u[(long)i * ld]);
    ++i;
  }
  if (S2 > 1) theta[(long)i * ldo] = sd * softplus(u[(long)i * ld]);
}

}  // namespace ci
