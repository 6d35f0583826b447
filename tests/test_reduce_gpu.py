"""GPU parity of the sample reductions (ci_reduce_inferences / ci_reduce_summary) against the
reference's NumPy semantics restated in oracle/sts_oracle.py (inferences.py:52-249, 285-408).
Tolerance: fp32 arithmetic on identical samples -> rtol 2e-5 (+ small atol for sums near 0)."""
import numpy as np
import pytest
import torch

from helpers import O

pytestmark = pytest.mark.gpu


def _case(N, T, Tf, nsamp, seed, mask_frac=0.0):
    rng = np.random.default_rng(seed)
    y_pre = rng.normal(100, 5, size=(N, T)).astype(np.float32)
    y_post = rng.normal(105, 5, size=(N, Tf)).astype(np.float32)
    if mask_frac > 0:
        m = rng.random(size=(N, Tf)) < mask_frac
        m[:, 0] = False
        y_post[m] = np.nan
    pre_sims = (y_pre[:, None, :] + rng.normal(0, 3, size=(N, nsamp, T))).astype(np.float32)
    post_sims = (100 + rng.normal(0, 3, size=(N, nsamp, Tf)).cumsum(axis=2)).astype(np.float32)
    pre_mean = pre_sims.mean(axis=1).astype(np.float32)
    post_mean = post_sims.mean(axis=1).astype(np.float32)
    return y_pre, y_post, pre_sims, pre_mean, post_sims, post_mean


@pytest.mark.parametrize('N,T,Tf,nsamp,alpha,mask_frac', [
    (3, 40, 12, 1000, 0.05, 0.0), (2, 17, 9, 100, 0.1, 0.3), (1, 70, 30, 1000, 0.05, 0.0),
    (2, 10, 5, 1500, 0.2, 0.2), (1, 12, 7, 3000, 0.05, 0.0), (1, 9, 4, 7, 0.5, 0.0),
    (2, 20, 130, 300, 0.05, 0.2),      # post period longer than the shared-memory running-sum tile (direct kernel)
    (3, 15, 96, 257, 0.05, 0.1)])      # ... exactly at its limit, a partial last 128-row tile
def test_reductions_match_numpy(N, T, Tf, nsamp, alpha, mask_frac):
    from tfcausalimpact_b200.engine import DeviceBatch
    y_pre, y_post, pre_sims, pre_mean, post_sims, post_mean = _case(N, T, Tf, nsamp, N * 7 + T, mask_frac)
    db = DeviceBatch(y_pre, None, Tf=Tf)
    dev = [torch.as_tensor(a).cuda() for a in (y_pre, y_post, pre_sims, pre_mean, post_sims, post_mean)]
    inf = db.reduce_inferences(*dev, alpha).cpu().numpy()
    summ = db.reduce_summary(dev[1], dev[5], dev[4], alpha).cpu().numpy()
    names = ['complete_preds_means', 'complete_preds_lower', 'complete_preds_upper', 'post_preds_means',
             'post_preds_lower', 'post_preds_upper', 'post_cum_y', 'post_cum_preds_means',
             'post_cum_preds_lower', 'post_cum_preds_upper', 'point_effects_means', 'point_effects_lower',
             'point_effects_upper', 'post_cum_effects_means', 'post_cum_effects_lower',
             'post_cum_effects_upper']
    for n in range(N):
        mask = ~np.isnan(y_post[n])
        ref = O.reduce_inferences(y_pre[n], y_post[n][mask], pre_mean[n], post_mean[n][mask], pre_sims[n],
                                  post_sims[n][:, mask], alpha)
        for c, name in enumerate(names):
            exp = ref[name]
            got = inf[n, c, :len(exp)]
            np.testing.assert_allclose(got, exp, rtol=2e-5, atol=2e-4, err_msg=name)
            assert np.all(inf[n, c, len(exp):] == 0.0)
        tab = O.reduce_summary(y_post[n][mask], post_mean[n][mask], post_sims[n][:, mask], alpha)
        np.testing.assert_allclose(summ[n, :20].reshape(10, 2), tab, rtol=3e-5, atol=1e-5)
        p = O.p_value(post_sims[n][:, mask], y_post[n][mask].sum())
        assert abs(summ[n, 20] - p) < 1e-7


def test_reference_known_answer_p_value():
    """tests/test_inferences.py:448-455 of the reference: p == 2/11 for its fixed simulation."""
    from tfcausalimpact_b200.engine import DeviceBatch
    sims = np.array([[1, 1, 1], [1, 1, 1], [2, 2, 2], [3, 3, 3], [4, 4, 4], [5, 5, 5], [6, 6, 6],
                     [7, 7, 7], [8, 8, 8], [9, 9, 9]], dtype=np.float32)[None]
    y_post = np.array([[3.0, 3.0, 3.1]], dtype=np.float32)         # sum 9.1 -> 2 rows below... see oracle
    db = DeviceBatch(np.zeros((1, 4), dtype=np.float32), None, Tf=3)
    summ = db.reduce_summary(torch.as_tensor(y_post).cuda(), torch.zeros(1, 3).cuda() + 1,
                             torch.as_tensor(sims).cuda(), 0.05).cpu().numpy()
    assert abs(summ[0, 20] - O.p_value(sims[0], y_post.sum())) < 1e-7


@pytest.mark.parametrize('nsamp', [1000, 1024, 1025, 2048, 33])
def test_percentiles_with_ties_and_negative_values(nsamp):
    """The radix-select path (<= 2048 draws) on columns full of ties, negative values and signed zeros."""
    from tfcausalimpact_b200.engine import DeviceBatch
    rng = np.random.default_rng(nsamp)
    N, T, Tf = 2, 11, 6
    y_pre = rng.normal(size=(N, T)).astype(np.float32)
    y_post = rng.normal(size=(N, Tf)).astype(np.float32)
    pre_sims = rng.integers(-3, 4, size=(N, nsamp, T)).astype(np.float32)          # heavy ties, both signs
    pre_sims[:, ::7, :] *= -0.0
    pre_sims[:, :, 0] = 5.0                                                          # constant column
    pre_sims[:, :, 1] = -np.abs(rng.normal(size=(N, nsamp))).astype(np.float32)      # all negative
    post_sims = (rng.normal(size=(N, nsamp, Tf)) * 1e-3).astype(np.float32)
    pre_mean = pre_sims.mean(axis=1).astype(np.float32)
    post_mean = post_sims.mean(axis=1).astype(np.float32)
    db = DeviceBatch(y_pre, None, Tf=Tf)
    dev = [torch.as_tensor(a).cuda() for a in (y_pre, y_post, pre_sims, pre_mean, post_sims, post_mean)]
    inf = db.reduce_inferences(*dev, 0.05).cpu().numpy()
    for n in range(N):
        lo, hi = O.percentiles(pre_sims[n], 0.05)
        np.testing.assert_allclose(inf[n, 1, :T], lo, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(inf[n, 2, :T], hi, rtol=1e-6, atol=1e-7)
        lo, hi = O.percentiles(post_sims[n], 0.05)
        np.testing.assert_allclose(inf[n, 4, :Tf], lo, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(inf[n, 5, :Tf], hi, rtol=1e-6, atol=1e-9)
