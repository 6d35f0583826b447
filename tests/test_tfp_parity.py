"""Kernel-level parity against the LITERAL reference arithmetic: `ci_kalman_logp_grad` vs TensorFlow Probability's
own `model.joint_log_prob` + `tf.GradientTape` on the model the reference's `build_default_model` makes
(/root/reference/causalimpact/model.py:239-321, 375-377), tolerance rtol 1e-4 (north_star).

TensorFlow Probability is not installable in this image (no wheel, no network; `/root/reference/setup.py:51-52`), so the
module is SKIPPED here and the pin is the statistical one of tests/test_reference_tables_gpu.py.  The moment the driver
provides TFP (site-packages or baseline/_ref) this test switches on, unchanged."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, 'baseline', '_ref')
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.insert(0, _REF)
HAVE_TFP = all(importlib.util.find_spec(m) is not None for m in ('tensorflow', 'tensorflow_probability'))

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HAVE_TFP, reason='tensorflow_probability not importable')]


def _tfp_model(y, X, S):
    import tensorflow as tf
    import tensorflow_probability as tfp
    tfd = tfp.distributions
    obs = tf.convert_to_tensor(y[:, None], dtype=tf.float32)
    T = y.shape[0]
    obs_sd = float(np.nanstd(y))
    # priors exactly as causalimpact/model.py:196-216,283-290 builds them (sqrt of an InverseGamma variance)
    def ig_sd_prior(sigma_guess, df=32.0):
        ss = df * sigma_guess ** 2 / 2.0
        return tfd.TransformedDistribution(tfd.InverseGamma(tf.constant(df / 2., tf.float32), tf.constant(ss, tf.float32)),
                                           tfp.bijectors.Invert(tfp.bijectors.Square()))
    comps = [tfp.sts.LocalLevel(level_scale_prior=ig_sd_prior(0.01), observed_time_series=obs)]
    if X is not None and X.shape[1] > 0:
        comps.append(tfp.sts.SparseLinearRegression(design_matrix=tf.convert_to_tensor(X, tf.float32),
                                                    weights_prior_scale=0.1))
    if S > 1:
        comps.append(tfp.sts.Seasonal(num_seasons=S, num_steps_per_season=1, observed_time_series=obs))
    model = tfp.sts.Sum(comps, observed_time_series=obs,
                        observation_noise_scale_prior=ig_sd_prior(obs_sd / 2.0))
    return model, obs


@pytest.mark.parametrize('S,k,T', [(1, 1, 70), (1, 3, 90), (7, 10, 365), (7, 50, 365)])
def test_logp_grad_matches_tfp(S, k, T):
    import tensorflow as tf
    import torch
    from helpers import synth_problem
    from tfcausalimpact_b200.engine import DeviceBatch
    Tf = 10
    y, _, X = synth_problem(1, T, Tf, k, S, 1, seed=11 + S + k)
    model, obs = _tfp_model(y[0], X[0], S)
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    rng = np.random.default_rng(0)
    G = 4
    u = (rng.normal(size=(db.P, G)) * 0.5).astype(np.float32)
    u[0] -= 1.0
    u[1] -= 2.0
    lp, g = db.logp_grad(torch.as_tensor(u).cuda())
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    # unconstrained -> constrained through TFP's own parameter bijectors, log-det included
    params = model.parameters
    sizes = [int(np.prod(p.prior.event_shape)) if p.prior.event_shape.rank else 1 for p in params]
    assert sum(sizes) == db.P, (sizes, db.P)
    for c in range(G):
        uu = [tf.Variable(u[o:o + n, c].reshape(p.prior.event_shape.as_list() or [])) for p, o, n in
              zip(params, np.cumsum([0] + sizes[:-1]), sizes)]
        with tf.GradientTape() as tape:
            theta = [p.bijector.forward(v) for p, v in zip(params, uu)]
            ldj = tf.add_n([tf.reduce_sum(p.bijector.forward_log_det_jacobian(v, event_ndims=len(v.shape)))
                            for p, v in zip(params, uu)])
            target = model.joint_distribution(obs).log_prob(*theta) + ldj
        grads = tape.gradient(target, uu)
        gt = np.concatenate([np.asarray(x).reshape(-1) for x in grads])
        np.testing.assert_allclose(lp[c], float(target), rtol=1e-4, atol=1e-3)
        gmax = np.abs(gt).max()
        assert np.all(np.abs(g[:, c] - gt) <= 1e-4 * np.abs(gt) + 1e-5 * gmax)
